// tools/gather_ceiling.cu -- what random 64-byte row gathers can reach on this GPU (measurement aid, not part of the product).
// n rows of 16 floats are read from a table of `rows` rows at precomputed uniformly random indices and written contiguously:
// one random 64-byte read + one streamed 64-byte write + the 4-byte index per row (4 lanes per row, one float4 each:
// fully coalesced stores, every load a whole 32-byte-sector pair).  `reads` = 2 reads a second random row per output row
// and adds it (the access pattern of bp_her_relabel: two random 64-byte reads per transition).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gather_ceiling tools/gather_ceiling.cu && tools/gather_ceiling
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

__global__ void gather(const float4* __restrict__ tab, const int* __restrict__ idx, const int* __restrict__ idx2, float4* __restrict__ out, long n, int reads) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long r = t >> 2;
    const int c = (int)(t & 3);
    if (r >= n) return;
    float4 v = __ldg(tab + (long)__ldg(idx + r) * 4 + c);
    if (reads == 2) {
        const float4 w = __ldg(tab + (long)__ldg(idx2 + r) * 4 + c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    __stcs(out + r * 4 + c, v);
}

__global__ void fill(float* p, long n, float v) { for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v; }
__global__ void sweep(const float* p, long n, float* o) { float s = 0; for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s += p[i]; if (s == 123.f) *o = s; }

int main(int argc, char** argv) {
    const double peak = argc > 1 ? atof(argv[1]) : 6536.7;
    const long rows = 20000L * 51, n = 1L << 20;   // the episode store of bench_her.py: 20000 episodes x 51 rows of 16 floats
    float4 *tab, *out; int *idx, *idx2; float *fl, *fl2;
    cudaMalloc(&tab, rows * 64); cudaMalloc(&out, n * 64); cudaMalloc(&idx, n * 4); cudaMalloc(&idx2, n * 4);
    cudaMalloc(&fl, 256L << 20); cudaMalloc(&fl2, 256L << 20);
    std::vector<int> h(n), h2(n);
    srand(1);
    for (long i = 0; i < n; ++i) { h[i] = (int)(((long)rand() * 32768 + rand()) % rows); h2[i] = (int)(((long)rand() * 32768 + rand()) % rows); }
    cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(idx2, h2.data(), n * 4, cudaMemcpyHostToDevice);
    fill<<<1184, 256>>>((float*)tab, rows * 16, 1.f);
    for (int reads = 1; reads <= 2; ++reads) {
        std::vector<float> ms;
        for (int it = 0; it < 23; ++it) {
            fill<<<1184, 256>>>(fl, 64L << 20, 0.f);            // L2 flush: 256 MB write, then a 256 MB read sweep (clean lines)
            sweep<<<1184, 256>>>(fl2, 64L << 20, fl);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            gather<<<(unsigned)((n * 4 + 255) / 256), 256>>>(tab, idx, idx2, out, n, reads);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float t; cudaEventElapsedTime(&t, a, b);
            if (it >= 3) ms.push_back(t);
        }
        std::sort(ms.begin(), ms.end());
        const double med = ms[ms.size() / 2], bytes = (double)n * (64.0 * reads + 64.0 + 4.0 * reads);
        printf("{\"metric\": \"random_64B_row_gather\", \"random_reads_per_row\": %d, \"ms\": %.4f, \"rows_per_s\": %.4g, \"GBps\": %.1f, \"frac_of_peak\": %.3f}\n",
               reads, med, n / (med * 1e-3), bytes / (med * 1e-3) / 1e9, bytes / (med * 1e-3) / 1e9 / peak);
    }
    return 0;
}
