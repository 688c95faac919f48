// af_internal.h -- what the translation units of the host runtime share (af_runtime.cu, af_multi.cu): the
// per-device library contexts, error plumbing and the internal form of a batch run.  Not part of the C ABI.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/audioflow_gpu.h"
#include "af_device.cuh"
#include "af_launch.h"
#include "af_plan.h"

namespace afrt {

using namespace af;

constexpr int MAX_DEV = AF_MAX_GPUS;

int fail(int code, const char *fmt, ...);

#define AF_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return ::afrt::fail(AF_ERR_CUDA, "CUDA error %s at %s:%d: %s", #expr, __FILE__, __LINE__, \
                                                   cudaGetErrorString(_e));                        \
    } while (0)

struct DevBuf {                       // grow-only device buffer
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct FracTable {                    // device copy of the f32 fractional offsets of one rate pair
    float *d = nullptr;
    size_t n = 0;
    ~FracTable() { if (d) cudaFree(d); }
};

struct RatePlan {                     // batch-path plan of one (in, out) pair, grown on demand
    RsRecurrence rec;                 // state after cum.size() - 1 chunks
    std::vector<uint64_t> cum{0};     // cum[c] = outputs after c chunks (table mode)
    std::vector<float> frac;          // host copy (table mode)
    std::shared_ptr<FracTable> dev;   // device copy covering frac.size() entries
};

// One context per GPU the process uses.  A thread works on ONE device at a time: the one it selected with af_init /
// af_init_multi, or the one the handle it passes was created on (DevScope).
struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;    // the stream the library enqueues on: its own (non-blocking) one, or the caller's (af_set_stream)
    cudaStream_t own = nullptr;
    cudaStream_t side = nullptr;      // second stream: the result gather runs here, off the compute stream
    FftTables *d_fft = nullptr;
    std::mutex mu;                    // guards the scratch buffers and the plan cache
    DevBuf scratch_in, scratch_out, scratch_aux, scratch_jobs;
    std::map<uint64_t, RatePlan> plans;
};

Context &cur_ctx();                   // context of the device this thread works on (valid after require_ctx / DevScope)
int cur_device();                     // ... and its index (-1: none yet)
int require_ctx();                    // selects (initialising on first use) the thread's device; AF_OK or an error
int init_device(int device);          // initialises `device` (idempotent) WITHOUT selecting it
Context *device_ctx(int device);      // nullptr when not initialised

// Makes `dev` the device of the calling thread (CUDA's and the library's) for the lifetime of the object and restores
// both afterwards: entry points that take a handle run on the device the handle was created on.
struct DevScope {
    int prev_lib, prev_cuda, dev;
    int rc;
    explicit DevScope(int dev_);
    ~DevScope();
};

#define AF_SCOPE(dev_)                      \
    ::afrt::DevScope scope_((dev_));        \
    if (scope_.rc != AF_OK) return scope_.rc

void count_launch(uint64_t n = 1);
const std::string &kernel_variant();

// device a device pointer lives on (-1: not a device pointer / unknown)
int device_of_pointer(const void *p);

// af_batch_run on an explicit stream with the VAD states written to `vad` (may differ from o->vad: the sharded
// batches point it into the gather buffer).  Never synchronises.
int batch_run_on(af_batch *b, const af_outputs *o, uint8_t *vad, uint64_t vad_stride, cudaStream_t st);
int batch_device(const af_batch *b);
int batch_check_outputs(const af_batch *b, const af_outputs *o);
const af_pipeline_config &pipeline_cfg(const af_pipeline *p);
int plan_stream_counts(const af_pipeline_config &cfg, const af_stream_desc &d, uint32_t *n_out, uint32_t *n_frames, uint32_t *n_vad);
uint64_t batch_max_vad(const af_batch *b);

// the communicator side (af_multi.cu)
void comm_shutdown_all();

}  // namespace afrt
