// stub (replaced below)
#include "ts_kernels.h"
namespace ts {
Tuning& tuning() { static Tuning t = {3, 48, 8, 2, 0}; return t; }
StagedPlan plan_staged(const Geo&, int, int, int, bool, const void*, const void*, const void*, int) { StagedPlan p{}; p.ok = false; return p; }
int staged_gather(const Geo&, const StagedPlan&, int, const void*, void*, unsigned long long, int, const void*, int, long long, cudaStream_t) { return TS_ERR_UNSUPPORTED; }
int staged_active_forward(const Geo&, const StagedPlan&, const void*, const void*, void*, cudaStream_t) { return TS_ERR_UNSUPPORTED; }
int staged_backward(const Geo&, const StagedPlan&, int, const void*, const void*, const void*, void*, void*, double*, cudaStream_t) { return TS_ERR_UNSUPPORTED; }
}
