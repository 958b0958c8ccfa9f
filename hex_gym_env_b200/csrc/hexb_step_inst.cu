// hexb_step_inst.cu - everything of libhexb.so that is templated on the board size, for ONE size: compiled once per
// N = 3..19 with -DHEXB_INST_N=n (hex_gym_env_b200/_native.py builds the 17 objects in parallel and links them with
// hexb_kernels.cu). nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo.
#ifndef HEXB_INST_N
#error "compile with -DHEXB_INST_N=<board size>"
#endif
#include "hexb_step.cuh"

#define HEXB_CAT_(a, b) a##b
#define HEXB_CAT(a, b) HEXB_CAT_(a, b)

template <int N>
__global__ void hexb_sample_kernel(View V, int view, const double *u, int32_t *out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < V.G) sample_at<N>(V, view, g, u, out);
}
template <int N>
__global__ void hexb_import_kernel(Params P, const int8_t *board_true, const int8_t *to_move, const uint8_t *import_mask) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < P.G) import_game<N>(P, g, board_true, to_move, import_mask);
}

int HEXB_CAT(hexb_launch_tile_, HEXB_INST_N)(const hexb_env *e, const Params &P, cudaStream_t s) { return launch_tile<HEXB_INST_N>(e, P, s); }

int HEXB_CAT(hexb_launch_sample_, HEXB_INST_N)(const View &V, int view, const double *u, int32_t *out, cudaStream_t s) {
    hexb_sample_kernel<HEXB_INST_N><<<(unsigned)((V.G + 127) / 128), 128, 0, s>>>(V, view, u, out);
    CK(cudaGetLastError());
    return HEXB_OK;
}

int HEXB_CAT(hexb_launch_import_, HEXB_INST_N)(const Params &P, const int8_t *board_true, const int8_t *to_move, const uint8_t *import_mask,
                                               cudaStream_t s) {
    hexb_import_kernel<HEXB_INST_N><<<(unsigned)((P.G + 127) / 128), 128, 0, s>>>(P, board_true, to_move, import_mask);
    CK(cudaGetLastError());
    return HEXB_OK;
}
