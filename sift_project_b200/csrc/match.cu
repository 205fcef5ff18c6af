// Matcher dispatch: tensor-core path for problems that fill tcgen05 tiles, SIMT dp4a otherwise.
// Both produce the same exact integer (best index, best d^2, second d^2) triples.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace sb {

cudaError_t match_init() { return match_tc_init(); }

// SIFT_B200_MATCH=simt|tc forces one path (tests exercise both on the same inputs).
bool match_uses_tensor_cores(int na, int nb) {
    const char* f = getenv("SIFT_B200_MATCH");
    if (f && f[0] == 's') return false;
    if (f && f[0] == 't') return na >= 1 && nb >= 1;
    return nb >= 256 && (long long)na * nb >= (1ll << 20);
}

cudaError_t launch_match(const uint8_t* a, int na, const uint8_t* b, int nb, int* best_idx, int* best_d2,
                         int* second_d2, const MatchScratch& ms, int sm_count, cudaStream_t s, int* launches) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    if (aligned && match_uses_tensor_cores(na, nb))
        return launch_match_tc(a, na, b, nb, best_idx, best_d2, second_d2, ms, sm_count, s, launches);
    return launch_match_simt(a, na, b, nb, best_idx, best_d2, second_d2, ms, sm_count, s, launches);
}

}  // namespace sb
