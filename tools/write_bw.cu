// Pure-write / mixed bandwidth probe: what can a store-dominated kernel reach on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k_write(int4 *out, const int4 *in, size_t n, int read_every) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int4 v = make_int4((int)i, 1, 2, 3);
    for (; i < n; i += stride) {
        if (read_every && (i / 32) % read_every == 0) { int4 r = in[i]; v.x ^= r.x; }
        if (MODE == 0) out[i] = v;
        else if (MODE == 1) __stcs(out + i, v);
        else if (MODE == 2) __stcg(out + i, v);
        else __stwt(out + i, v);
    }
}

// one warp writes CHUNK contiguous bytes per iteration with 8-byte stores (like other_positions)
__global__ void k_write8(int2 *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) __stcs(out + i, make_int2((int)i, 7));
}

template <typename F>
float timeit(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t bytes = (size_t)2304 << 20;  // 2.25 GiB, > L2
    int4 *out, *in;
    cudaMalloc(&out, bytes); cudaMalloc(&in, bytes);
    cudaMemset(in, 1, bytes);
    const size_t n = bytes / 16;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int bpsm : {2, 4, 8, 16, 32}) {
        const int grid = sms * bpsm;
        float t0 = timeit([&] { k_write<0><<<grid, 256>>>(out, in, n, 0); });
        float t1 = timeit([&] { k_write<1><<<grid, 256>>>(out, in, n, 0); });
        float t2 = timeit([&] { k_write<2><<<grid, 256>>>(out, in, n, 0); });
        float t3 = timeit([&] { k_write<3><<<grid, 256>>>(out, in, n, 0); });
        float t8 = timeit([&] { k_write8<<<grid, 256>>>((int2 *)out, bytes / 8); });
        float tm = timeit([&] { k_write<1><<<grid, 256>>>(out, in, n, 14); });  // ~7% reads
        printf("{\"blocks_per_sm\": %d, \"st_GBs\": %.0f, \"st_cs_GBs\": %.0f, \"st_cg_GBs\": %.0f, \"st_wt_GBs\": %.0f, "
               "\"st_cs_8B_GBs\": %.0f, \"st_cs_plus_7pct_reads_GBs\": %.0f}\n",
               bpsm, bytes / t0 / 1e6, bytes / t1 / 1e6, bytes / t2 / 1e6, bytes / t3 / 1e6, bytes / t8 / 1e6,
               bytes * (1 + 1.0 / 14) / tm / 1e6);
    }
    float tc = timeit([&] { cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice); });
    float ts = timeit([&] { cudaMemsetAsync(out, 3, bytes); });
    printf("{\"memcpy_d2d_GBs_rw\": %.0f, \"memset_GBs\": %.0f}\n", 2.0 * bytes / tc / 1e6, bytes / ts / 1e6);
    return 0;
}
