//! Golden vectors from rubato 0.16.2 `FastFixedIn<f32>` / `PolynomialDegree::Cubic`, driven exactly as
//! `AudioResampler::process` drives it (src-tauri/src/modules/audio/resampler.rs:71-93: one 128-frame chunk per call,
//! one channel) with the INTENDED constructor arguments (ratio = out / in, max relative ratio 1.0, chunk 128, 1 channel;
//! the literal call at resampler.rs:43-49 passes them in the wrong order -- its outcome is recorded below as well).
//!
//! Input: the deterministic sequence tests/test_oracle.py regenerates (splitmix64 -> 24-bit mantissa in [-1, 1)).
//! Output: JSON on stdout; every sample as its u32 bit pattern so that the comparison is bit-exact.
use rubato::{FastFixedIn, PolynomialDegree, Resampler};

fn splitmix64(state: &mut u64) -> u64 {
    *state = state.wrapping_add(0x9E3779B97F4A7C15);
    let mut z = *state;
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}

fn input(rate: u32, n: usize) -> Vec<f32> {
    let mut s: u64 = 0xA0D10F10u64 + rate as u64;
    (0..n).map(|_| ((splitmix64(&mut s) >> 40) as f32) / 8388608.0 - 1.0).collect()
}

fn main() {
    const CHUNK: usize = 128;
    const N_CHUNKS: usize = 40;
    let rates = [48000u32, 44100, 32000, 22050, 8000];
    println!("{{\n \"rubato\": \"0.16.2\", \"chunk\": {}, \"n_chunks\": {}, \"input\": \"splitmix64(0xA0D10F10 + rate) >> 40, / 2^23 - 1\",", CHUNK, N_CHUNKS);
    // the literal constructor call of the reference (resampler.rs:43-49): does it build, and what does process() say?
    let literal = FastFixedIn::<f32>::new(48000.0, 16000.0, PolynomialDegree::Cubic, 128, 128);
    match literal {
        Ok(mut r) => {
            let x = vec![0.0f32; CHUNK];
            let res = r.process(&[x], None);
            println!(" \"literal_reference_call\": {{\"constructs\": true, \"process\": {:?}}},", res.map(|v| v.len()).map_err(|e| e.to_string()));
        }
        Err(e) => println!(" \"literal_reference_call\": {{\"constructs\": false, \"error\": {:?}}},", e.to_string()),
    }
    println!(" \"rates\": {{");
    for (ri, &rate) in rates.iter().enumerate() {
        let x = input(rate, CHUNK * N_CHUNKS);
        let mut r = FastFixedIn::<f32>::new(16000.0 / rate as f64, 1.0, PolynomialDegree::Cubic, CHUNK, 1).expect("FastFixedIn::new");
        print!("  \"{}\": [", rate);
        for c in 0..N_CHUNKS {
            let out = r.process(&[x[c * CHUNK..(c + 1) * CHUNK].to_vec()], None).expect("process");
            let bits: Vec<String> = out[0].iter().map(|v| v.to_bits().to_string()).collect();
            print!("{}[{}]", if c == 0 { "" } else { "," }, bits.join(","));
        }
        println!("]{}", if ri + 1 == rates.len() { "" } else { "," });
    }
    println!(" }}\n}}");
}
