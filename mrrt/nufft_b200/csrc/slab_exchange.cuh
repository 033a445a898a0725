// Exchange of grid rows between the plane stage and the axis-3 stage of the slab-distributed
// transforms (SlabShardedNufft), written against PEER memory: every rank's slab grid is mapped
// into every other rank's address space (torch symmetric memory over NVLink / NVSwitch), so the
// all-to-all, its pack / unpack passes and the halo summation are ONE kernel each:
//
//   forward  slab_scatter_kernel: rank r stores the rows of its planes [z0, z0 + nz) straight
//            into the slab grids of the ranks that hold those rows (a row in a halo goes to two
//            or more ranks);
//   adjoint  slab_gather_kernel: rank r reads its planes out of every slab grid and writes, per
//            grid row, the SUM over the slabs that hold the row (its owner plus the halos of
//            the slabs before it) -- the halo reduction rides along, in a fixed order.
//
// One warp moves one grid row (K1 complex values) with 16-byte accesses.  The reference has no
// counterpart (single device).
#pragma once
#include "common.cuh"

namespace b2n {

constexpr int kMaxPeers = 16;
constexpr int kExchUnroll = 4;      // 16-byte accesses per lane in flight

struct SlabPeers {
    void* grid[kMaxPeers];     // slab grid of every rank, [K3][nrows][K1] complex, peer-mapped
    int row0[kMaxPeers];       // global row (axis 2) of local row 0
    int nrows[kMaxPeers];      // rows held (owned rows + halo)
    int world;
};

template <typename V>
__device__ __forceinline__ V vadd(V a, V b);
template <> __device__ __forceinline__ float4 vadd(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
template <> __device__ __forceinline__ double2 vadd(double2 a, double2 b) {
    return make_double2(a.x + b.x, a.y + b.y);
}

// planes: [nz][K2][K1] complex (this rank's planes after the in-plane FFT).  V = 16-byte vector
// of the real type; vpr = vectors per row (K1 * sizeof(complex) / 16).
template <typename V>
__global__ void __launch_bounds__(256)
slab_scatter_kernel(SlabPeers P, const V* __restrict__ planes, int nz, int z0, int K2, int vpr) {
    int total_rows = 0;
    for (int s = 0; s < P.world; s++) total_rows += P.nrows[s];
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t items = (int64_t)nz * total_rows;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += nwarp) {
        const int z = (int)(it / total_rows);
        int r = (int)(it - (int64_t)z * total_rows);
        int s = 0;
        while (r >= P.nrows[s]) { r -= P.nrows[s]; s++; }
        int k2 = P.row0[s] + r;
        if (k2 >= K2) k2 -= K2;
        const V* __restrict__ src = planes + ((int64_t)z * K2 + k2) * vpr;
        V* __restrict__ dst = (V*)P.grid[s] + ((int64_t)(z0 + z) * P.nrows[s] + r) * vpr;
        // four independent 16-byte loads per lane in flight before the peer stores
        for (int e0 = lane; e0 < vpr; e0 += 32 * kExchUnroll) {
            V v[kExchUnroll];
#pragma unroll
            for (int u = 0; u < kExchUnroll; u++)
                if (e0 + 32 * u < vpr) v[u] = src[e0 + 32 * u];
#pragma unroll
            for (int u = 0; u < kExchUnroll; u++)
                if (e0 + 32 * u < vpr) dst[e0 + 32 * u] = v[u];
        }
    }
}

template <typename V>
__global__ void __launch_bounds__(256)
slab_gather_kernel(SlabPeers P, V* __restrict__ planes, int nz, int z0, int K2, int vpr) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t items = (int64_t)nz * K2;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += nwarp) {
        const int z = (int)(it / K2);
        const int k2 = (int)(it - (int64_t)z * K2);
        V* __restrict__ dst = planes + it * vpr;
        // the slabs holding this row are the same for the whole row: the peer loads of a chunk
        // (kExchUnroll x 16 bytes per lane) are independent and all in flight together
        for (int e0 = lane; e0 < vpr; e0 += 32 * kExchUnroll) {
            V acc[kExchUnroll];
            bool have = false;
            for (int s = 0; s < P.world; s++) {
                int r = k2 - P.row0[s];
                if (r < 0) r += K2;
                if (r < P.nrows[s]) {
                    const V* __restrict__ q = (const V*)P.grid[s] + ((int64_t)(z0 + z) * P.nrows[s] + r) * vpr;
#pragma unroll
                    for (int u = 0; u < kExchUnroll; u++) {
                        if (e0 + 32 * u < vpr) {
                            const V v = q[e0 + 32 * u];
                            acc[u] = have ? vadd<V>(acc[u], v) : v;      // fixed order: slab index
                        }
                    }
                    have = true;
                }
            }
#pragma unroll
            for (int u = 0; u < kExchUnroll; u++)
                if (e0 + 32 * u < vpr) dst[e0 + 32 * u] = acc[u];      // every row has an owner: have is true
        }
    }
}

}  // namespace b2n
