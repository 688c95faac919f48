//! Raw bindings to include/audioflow_gpu.h (what `bindgen` would emit, trimmed to the compat API).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct af_resampler { _p: [u8; 0] }
#[repr(C)]
pub struct af_batch_resampler { _p: [u8; 0] }
#[repr(C)]
pub struct af_vad { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct af_vad_config {
    pub threshold_db: f32,
    pub smoothing_factor: f32,
    pub silence_timeout_frames: u64,
    pub min_speech_frames: u64,
}

pub const AF_OK: c_int = 0;

#[repr(C)] pub struct af_ring { _private: [u8; 0] }
pub const AF_RING_EMPTY: c_int = -1;

extern "C" {
    pub fn af_ring_create(capacity_samples: usize, out: *mut *mut af_ring) -> c_int;
    pub fn af_ring_destroy(r: *mut af_ring);
    pub fn af_ring_write(r: *mut af_ring, data: *const f32, n: usize) -> usize;
    pub fn af_ring_read(r: *mut af_ring, out: *mut f32, size: usize, n_read: *mut usize) -> c_int;
    pub fn af_ring_available(r: *const af_ring) -> usize;
    pub fn af_ring_clear(r: *mut af_ring);

    pub fn af_init(device: c_int) -> c_int;
    pub fn af_last_error(buf: *mut c_char, cap: usize) -> usize;

    pub fn af_to_mono(samples: *const f32, n_samples: usize, channels: u16, out: *mut f32, out_cap: usize,
                      n_out: *mut usize) -> c_int;

    pub fn af_resampler_create(input_rate: u32, output_rate: u32, out: *mut *mut af_resampler) -> c_int;
    pub fn af_resampler_destroy(r: *mut af_resampler);
    pub fn af_resampler_process(r: *mut af_resampler, input: *const f32, n: usize, out: *mut f32, out_cap: usize,
                                n_out: *mut usize) -> c_int;
    pub fn af_resample_max_output(input_rate: u32, output_rate: u32, n_in: usize) -> usize;

    pub fn af_batch_resampler_create(input_rate: u32, output_rate: u32, out: *mut *mut af_batch_resampler) -> c_int;
    pub fn af_batch_resampler_destroy(b: *mut af_batch_resampler);
    pub fn af_batch_resampler_process(b: *mut af_batch_resampler, input: *const f32, n: usize, out: *mut f32,
                                      out_cap: usize, n_out: *mut usize) -> c_int;
    pub fn af_batch_resampler_flush(b: *mut af_batch_resampler, out: *mut f32, out_cap: usize, n_out: *mut usize) -> c_int;

    pub fn af_vad_create(cfg: *const af_vad_config, out: *mut *mut af_vad) -> c_int;
    pub fn af_vad_destroy(v: *mut af_vad);
    pub fn af_vad_detect(v: *mut af_vad, frame: *const f32, n: usize, state: *mut u8) -> c_int;
    pub fn af_vad_reset(v: *mut af_vad) -> c_int;
    pub fn af_vad_state(v: *const af_vad) -> c_int;
    pub fn af_vad_energy_db(v: *const af_vad) -> f32;
    pub fn af_vad_is_speaking(v: *const af_vad) -> c_int;
    pub fn af_vad_speech_frame_count(v: *const af_vad) -> u64;
}
