// Loader kernels (reference wav.py:64-91,108-156): placeholder until K1/K2 land.
#include "sb_internal.h"
using namespace sb;
extern "C" {
int sb_load_pcm(const void*, int64_t, int, int, int, int, int64_t, int64_t, sb_stream**) {
    SB_FAIL(SB_ESTATE, "sb_load_pcm: not built yet");
}
int sb_normalise(const sb_stream*, int, sb_stream**, float*, float*) {
    SB_FAIL(SB_ESTATE, "sb_normalise: not built yet");
}
}
