// Read-pattern microbenchmark for the posteriors x (B,T,ld): strided 2 KB row segments (current tiling) vs a
// tile-major layout where each CTA's stream is contiguous.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int ROWS>
__global__ void __launch_bounds__(128, 4) k_read(const float* __restrict__ x, int B, int T, long long ld, int nvt, int tile_major,
                                                 int ntiles, float* out) {
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int vt = tile % nvt, b = tile / nvt;
        for (int t0 = 0; t0 < T; t0 += ROWS) {
            float4 v[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int t = min(t0 + r, T - 1);
                const float* p = tile_major ? x + (((size_t)(b * nvt + vt) * T + t) * 512 + threadIdx.x * 4)
                                            : x + ((size_t)(b * T + t) * ld + vt * 512 + threadIdx.x * 4);
                v[r] = __ldcs(reinterpret_cast<const float4*>(p));
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc += v[r].x + v[r].y + v[r].z + v[r].w;
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}

extern "C" int membw3_run(const float* x, int B, int T, long long ld, int nvt, int tile_major, int rows, int ctas_per_sm, float* out, float* ms_out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int ntiles = B * nvt, grid = 148 * ctas_per_sm;
    auto launch = [&] {
        if (rows == 8) k_read<8><<<grid, 128>>>(x, B, T, ld, nvt, tile_major, ntiles, out);
        else if (rows == 16) k_read<16><<<grid, 128>>>(x, B, T, ld, nvt, tile_major, ntiles, out);
        else k_read<4><<<grid, 128>>>(x, B, T, ld, nvt, tile_major, ntiles, out);
    };
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); *ms_out = ms / 5;
    return cudaGetLastError();
}
