// Microbenchmarks that decide the adjoint flush strategy on B200:
//  1. coalesced REDG.F32x2: a warp adds 32 consecutive complex cells
//  2. strided   REDG.F32x2: each lane its own 32-byte sector (rows K1 apart)
//  3. cp.reduce.async.bulk (smem -> global add.f32) of 176-byte rows, one op per lane
//  4. smem ATOMS.ADD (u32) spread / float CAS for reference
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void red_coalesced(float2* g, int64_t ncell, int iters) {
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int it = 0; it < iters; it++) {
        int64_t base = ((warp + (int64_t)it * nwarp) * 2654435761ull) % (ncell / 32);
        atomicAdd(&g[base * 32 + lane], make_float2(1.f, 2.f));
    }
}
__global__ void red_strided(float2* g, int64_t ncell, int K1, int iters) {
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int it = 0; it < iters; it++) {
        int64_t base = ((warp + (int64_t)it * nwarp) * 2654435761ull) % (ncell - (int64_t)40 * K1);
        atomicAdd(&g[base + (int64_t)lane * K1], make_float2(1.f, 2.f));
    }
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each warp owns a [36][22] complex buffer; per iteration lanes issue one bulk reduce per row
__global__ void __launch_bounds__(128) red_bulk(float2* g, int64_t ncell, int K1, int K2, int iters, int ncells_row) {
    __shared__ __align__(128) float2 buf[4][36][22];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int r = lane; r < 36; r += 32)
        for (int c = 0; c < 22; c++) buf[wib][r][c] = make_float2(1.f, 2.f);
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int64_t nseg = ncell / 16 - (int64_t)8 * K1 * K2 / 16;
    for (int it = 0; it < iters; it++) {
        int64_t seg = ((warp + (int64_t)it * nwarp) * 2654435761ull) % nseg;
        int64_t base = seg * 16;
        for (int r = lane; r < 36; r += 32) {
            int64_t row = base + (int64_t)(r % 6) * K1 + (int64_t)(r / 6) * K1 * K2;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                         :: "l"(g + row), "r"(smem_u32(&buf[wib][r][0])), "r"(ncells_row * 8) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void smem_atom_u32(unsigned* out, int iters) {
    __shared__ unsigned s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = 0;
    __syncthreads();
    unsigned idx = threadIdx.x * 37u;
    for (int it = 0; it < iters; it++) { atomicAdd(&s[idx & 4095], 1u); idx += 1031u; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}
__global__ void smem_atom_f32(float* out, int iters) {
    __shared__ float s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = 0;
    __syncthreads();
    unsigned idx = threadIdx.x * 37u;
    for (int it = 0; it < iters; it++) { atomicAdd(&s[idx & 4095], 1.f); idx += 1031u; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[5];
}

int main() {
    const int K1 = 384, K2 = 384, K3 = 384;
    const int64_t ncell = (int64_t)K1 * K2 * K3;
    float2* g; CK(cudaMalloc(&g, ncell * 8)); CK(cudaMemset(g, 0, ncell * 8));
    unsigned* o; CK(cudaMalloc(&o, 1 << 20));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    const int blocks = 148 * 16, iters = 2000;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a); red_coalesced<<<blocks, 128>>>(g, ncell, iters); cudaEventRecord(b); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, a, b);
        double n = (double)blocks * 4 * iters * 32;
        printf("coalesced REDG.F32x2: %.3f ms, %.1f G cell-adds/s\n", ms, n / ms / 1e6);
        cudaEventRecord(a); red_strided<<<blocks, 128>>>(g, ncell, K1, iters); cudaEventRecord(b); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, a, b);
        printf("strided   REDG.F32x2: %.3f ms, %.1f G cell-adds/s\n", ms, n / ms / 1e6);
        for (int nc = 6; nc <= 22; nc += 8) {
            cudaEventRecord(a); red_bulk<<<blocks, 128>>>(g, ncell, K1, K2, iters / 4, nc); cudaEventRecord(b); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms, a, b);
            double ops = (double)blocks * 4 * (iters / 4) * 36;
            printf("bulk reduce rows of %2d cells: %.3f ms, %.1f M ops/s, %.1f G cell-adds/s\n", nc, ms, ops / ms / 1e3, ops * nc / ms / 1e6);
        }
        cudaEventRecord(a); smem_atom_u32<<<148 * 8, 256>>>(o, 4000); cudaEventRecord(b); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, a, b);
        printf("smem ATOMS.ADD.u32 spread: %.3f ms, %.1f G atom/s (%.2f cyc/warp-instr/SM @1.9GHz)\n", ms, 148.0 * 8 * 256 * 4000 / ms / 1e6,
               ms * 1e-3 * 1.9e9 / (8.0 * 8 * 4000));
        cudaEventRecord(a); smem_atom_f32<<<148 * 8, 256>>>((float*)o, 4000); cudaEventRecord(b); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, a, b);
        printf("smem atomicAdd(float) CAS spread: %.3f ms, %.1f G atom/s (%.2f cyc/warp-instr/SM)\n", ms, 148.0 * 8 * 256 * 4000 / ms / 1e6,
               ms * 1e-3 * 1.9e9 / (8.0 * 8 * 4000));
    }
    return 0;
}
