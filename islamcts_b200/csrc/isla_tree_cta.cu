// CTA-per-tree general engine (vanilla / APW / SPW / DPW, leaf-parallel rollout batches).
// Placeholder until the kernel lands: every entry point fails loudly (no CPU fallback).
#include "isla_internal.h"

namespace isla {
static int nope() { set_error("CTA-per-tree engine not built yet"); return ISLA_ERR_INVALID; }
int cta_fill_layout(const isla_config&, isla_layout&) { return nope(); }
int cta_set_roots(isla_engine*, const double*, cudaStream_t) { return nope(); }
int cta_run(isla_engine*, int, long long, const uint8_t*, const double*, long long, cudaStream_t) { return nope(); }
int cta_root_stats(isla_engine*, int64_t*, double*, cudaStream_t) { return nope(); }
int cta_best(isla_engine*, int, int32_t*, cudaStream_t) { return nope(); }
}  // namespace isla
