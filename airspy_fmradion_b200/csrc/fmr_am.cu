// fmr_am.cu — AM handle (AmDecoder::process, AmDecode.cpp:96-218). Placeholder until the
// AM kernels land: every entry point reports FMR_ERR_UNSUPPORTED.
#include "fmr_host.cuh"
using namespace fmr;
struct fmr_am { int dummy; };
extern "C" fmr_status fmr_am_create(const fmr_am_config *, fmr_am **) { return fail(FMR_ERR_UNSUPPORTED, "AM path not built yet"); }
extern "C" void fmr_am_destroy(fmr_am *) {}
extern "C" fmr_status fmr_am_process_host(fmr_am *, const float *, size_t, const uint32_t *, uint32_t, double *, size_t, uint32_t *) { return fail(FMR_ERR_UNSUPPORTED, "AM path not built yet"); }
extern "C" fmr_status fmr_am_process_device(fmr_am *, const float *, size_t, const uint32_t *, uint32_t, double *, size_t, uint32_t *, void *) { return fail(FMR_ERR_UNSUPPORTED, "AM path not built yet"); }
extern "C" fmr_status fmr_am_query_output(fmr_am *, const uint32_t *, uint32_t, uint64_t *, uint32_t *) { return fail(FMR_ERR_UNSUPPORTED, "AM path not built yet"); }
extern "C" fmr_status fmr_am_stats(fmr_am *, uint32_t, fmr_am_stats_t *) { return fail(FMR_ERR_UNSUPPORTED, "AM path not built yet"); }
extern "C" uint32_t fmr_am_last_launches(fmr_am *) { return 0; }
extern "C" fmr_status fmr_am_set_profiling(fmr_am *, int) { return fail(FMR_ERR_UNSUPPORTED, "AM path not built yet"); }
extern "C" fmr_status fmr_am_stage_times(fmr_am *, float *, const char **, uint32_t, uint32_t *) { return fail(FMR_ERR_UNSUPPORTED, "AM path not built yet"); }
