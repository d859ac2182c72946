// Write-pattern sweep for the state tensor r (T,2,BW,ld): which CTA tiling / ordering reaches linear-stream bandwidth?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// CTA = (b, v-tile, hyp group).  NT threads, each writes NV float4 per row at columns tid*4 + j*NT*4.
// rows per frame: HW hyps x 2 planes.  order: 0 = g fastest, then vt, then b; 1 = vt fastest, then g, then b; 2 = b fastest
template <int NV, int CS>
__global__ void k_pat(float* r, int B, int W, int T, int V, int ld, int HW, int G, int nvt, int order) {
    int idx = blockIdx.x, g, vt, b;
    if (order == 0) { g = idx % G; idx /= G; vt = idx % nvt; b = idx / nvt; }
    else if (order == 1) { vt = idx % nvt; idx /= nvt; g = idx % G; b = idx / G; }
    else { b = idx % B; idx /= B; vt = idx % nvt; g = idx / nvt; }
    const int NT = blockDim.x;
    size_t BW = (size_t)B * W;
    size_t plane = BW * ld, frame = 2 * plane;
    int nh = min(HW, W - g * HW);
    float* base = r + ((size_t)(b * W + g * HW)) * ld + (size_t)vt * NT * NV * 4 + threadIdx.x * 4;
    float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int t = 0; t < T; ++t) {
        float* rp = base + (size_t)t * frame;
        for (int hh = 0; hh < nh; ++hh) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                int v = vt * NT * NV * 4 + threadIdx.x * 4 + j * NT * 4;
                if (v < V) {
                    if (CS) { __stcs((float4*)(rp + (size_t)hh * ld + j * NT * 4), val); __stcs((float4*)(rp + (size_t)hh * ld + j * NT * 4 + plane), val); }
                    else { *(float4*)(rp + (size_t)hh * ld + j * NT * 4) = val; *(float4*)(rp + (size_t)hh * ld + j * NT * 4 + plane) = val; }
                }
            }
        }
    }
}

extern "C" int membw2_run(float* buf, int B, int W, int T, int V, int ld, int NT, int NV, int HW, int order, int cs, float* out_ms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int G = (W + HW - 1) / HW;
    int nvt = (V + NT * NV * 4 - 1) / (NT * NV * 4);
    int grid = B * nvt * G;
    auto launch = [&] {
        #define L(nv) if (NV == nv) { if (cs) k_pat<nv,1><<<grid, NT>>>(buf, B, W, T, V, ld, HW, G, nvt, order); else k_pat<nv,0><<<grid, NT>>>(buf, B, W, T, V, ld, HW, G, nvt, order); }
        L(1) L(2) L(3) L(5) L(10)
    };
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 3; ++i) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); *out_ms = ms / 3;
    return cudaGetLastError();
}
