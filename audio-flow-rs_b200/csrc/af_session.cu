// af_session.cu -- streaming sessions (BASELINE config 5): persistent per-stream state on the device.
#include "../../include/audioflow_gpu.h"

extern "C" {

// TODO(round 1, later today): real implementation; until then the calls fail loudly.
int af_session_fail(void);

AF_API int af_session_create(af_pipeline *, size_t, uint32_t, uint16_t, uint16_t, uint32_t, af_session **out)
{
    if (out) *out = nullptr;
    return af_session_fail();
}
AF_API void af_session_destroy(af_session *) {}
AF_API int af_session_push(af_session *, const void *, uint64_t, uint32_t, int, const af_outputs *, uint32_t *,
                           uint32_t *, uint32_t *)
{
    return af_session_fail();
}
AF_API int af_session_reset(af_session *) { return af_session_fail(); }
}
