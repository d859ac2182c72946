// Write-bandwidth microbenchmarks (diagnostic tool, not product code): how fast can B200 absorb a pure
// float4 store stream shaped like the state tensor r of the CTC prefix scorer?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__global__ void k_store(float4* p, size_t n4, float v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    float4 val = make_float4(v, v, v, v);
    for (; i < n4; i += stride) {
        if (MODE == 0) p[i] = val;
        else if (MODE == 1) __stcs(p + i, val);
        else if (MODE == 2) __stwt(p + i, val);
        else __stcg(p + i, val);
    }
}

// scorer-like pattern: CTA owns (b, vtile, g); per frame writes HW*2 rows of 2 KB at stride V*4, planes BW*V apart
template <int MODE>
__global__ void k_store_pattern(float* r, int B, int W, int T, int V, int G, int HW, int nvt) {
    int idx = blockIdx.x;
    int g = idx % G; idx /= G;
    int vt = idx % nvt; int b = idx / nvt;
    int v0 = vt * 512 + threadIdx.x * 4;
    if (v0 >= V) return;
    size_t BW = (size_t)B * W;
    size_t plane = BW * V, frame = 2 * plane;
    float* base = r + ((size_t)(b * W + g * HW)) * V + v0;
    float4 val = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int t = 0; t < T; ++t) {
        float* rp = base + (size_t)t * frame;
        for (int hh = 0; hh < HW; ++hh) {
            if (MODE == 0) { *(float4*)(rp + (size_t)hh * V) = val; *(float4*)(rp + (size_t)hh * V + plane) = val; }
            else { __stcs((float4*)(rp + (size_t)hh * V), val); __stcs((float4*)(rp + (size_t)hh * V + plane), val); }
        }
    }
}

extern "C" int membw_run(float* buf, size_t bytes, int B, int W, int T, int V, float* out_ms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t n4 = bytes / 16;
    int k = 0;
    auto time = [&](auto f) { f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int i = 0; i < 3; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); out_ms[k++] = ms / 3; };
    time([&] { k_store<0><<<148 * 16, 256>>>((float4*)buf, n4, 1.f); });
    time([&] { k_store<1><<<148 * 16, 256>>>((float4*)buf, n4, 1.f); });
    time([&] { k_store<2><<<148 * 16, 256>>>((float4*)buf, n4, 1.f); });
    time([&] { k_store<3><<<148 * 16, 256>>>((float4*)buf, n4, 1.f); });
    time([&] { k_store<0><<<148 * 8, 1024>>>((float4*)buf, n4, 1.f); });
    time([&] { cudaMemsetAsync(buf, 0, bytes); });
    int G = 2, HW = 5, nvt = (V + 511) / 512;
    time([&] { k_store_pattern<0><<<B * nvt * G, 128>>>(buf, B, W, T, V, G, HW, nvt); });
    time([&] { k_store_pattern<1><<<B * nvt * G, 128>>>(buf, B, W, T, V, G, HW, nvt); });
    return cudaGetLastError();
}
