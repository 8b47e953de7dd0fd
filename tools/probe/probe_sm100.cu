// Micro-probes that decide the conv kernel's tiling (round 2). Stand-alone: nvcc -> one executable, run under gpurun.
//   1. tcgen05.mma issue cost vs (M, N) with both operands in shared memory, and with A in TMEM
//   2. per-SM TMA ingest from L2: tiled vs im2col, ring depth, grid size, cluster multicast
//   3. correctness of a ROW-SHIFTED shared-memory descriptor on a SWIZZLE_128B K-major tile
//      (the "slab" formulation of a 3x3 convolution: nine taps = nine row shifts of one staged tile)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_sm100 probe_sm100.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../anakin_b200/csrc/ptx.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_tiled;
static PFN_encodeIm2col g_im2col;

// ------------------------------------------------------------------------------------------------ 1. MMA cost
// mode 0: A,B from smem (SS); mode 1: A from TMEM (TS). kind: 0 i8, 1 f16, 2 tf32.
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if constexpr (KIND == KIND_I8) {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d), "r"(tmem_a),
                     "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d), "r"(tmem_a),
                     "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_cost_kernel(uint32_t idesc, int n_mma, int mode, int n_acc, long long* out, int a_shift_rows = 0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    // operand ring: 4 stages of (A 16 KiB + B 32 KiB), zero-filled
    for (int i = threadIdx.x; i < 4 * 49152 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    // warp 1, one elected lane, everything derived from warp-uniform values: the compiler keeps the descriptors in
    // uniform registers and emits back-to-back UTC*MMA (a divergent `threadIdx.x == 0` makes it wrap every MMA in an
    // ELECT / R2UR waterfall loop, which costs ~120 clk per MMA by itself)
    if (warp == 1) {
        if (elect_one()) {
            const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SW128
            const uint32_t base16 = smem_u32(smem) >> 4;
            for (int i = 0; i < 8; ++i) {
                const uint32_t a_lo = ((base16 + 2 * (i & 3)) & 0x3FFF) | (1u << 16), b_lo = ((base16 + 1024 + 2 * (i & 3)) & 0x3FFF) | (1u << 16);
                tc_mma_lohi<KIND>(tmem, a_lo, hi, b_lo, hi, idesc, i > 0);
            }
            tc_commit(&bar);
            mbar_wait(&bar, 0);
            tc_fence_after();
            const uint32_t d1 = tmem + (n_acc > 1 ? 256u : 0u);
            const uint32_t a_base = ((base16 + 8u * a_shift_rows) & 0x3FFF) | (1u << 16), b_base = ((base16 + 1024) & 0x3FFF) | (1u << 16);
            const long long t0 = clock64();
            if (mode == 0) {
                for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        // stage (u >> 2) of the ring, 32-byte k slice (u & 3)
                        const uint32_t off = (u >> 2) * (49152u >> 4) + 2u * (u & 3);
                        tc_mma_lohi<KIND>((u & 1) ? d1 : tmem, a_base + off, hi, b_base + off, hi, idesc, 1);
                    }
                }
            } else {
                for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t off = (u >> 2) * (49152u >> 4) + 2u * (u & 3);
                        mma_ts<KIND>(tmem, tmem + 448 + 8 * (u & 3), (static_cast<uint64_t>(hi) << 32) | (b_base + off), idesc, 1);
                    }
                }
            }
            const long long t1 = clock64();
            tc_commit(&bar);
            mbar_wait(&bar, 1);
            const long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

static uint32_t idesc_for(int kind, int M, int N) {
    // c_fmt: s32=2 for i8, f32=1 otherwise; a/b fmt: i8: s8=1; f16: 0; tf32: 2
    if (kind == KIND_I8) return make_idesc(2, 1, 1, M, N);
    if (kind == KIND_F16) return make_idesc(1, 0, 0, M, N);
    return make_idesc(1, 2, 2, M, N);
}

static void run_mma_cost() {
    long long* d_out;
    CK(cudaMalloc(&d_out, 16));
    const int smem = 4 * 49152 + 1024;
    CK(cudaFuncSetAttribute(mma_cost_kernel<KIND_I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mma_cost_kernel<KIND_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(mma_cost_kernel<KIND_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    printf("== 1. tcgen05.mma cost (clk per MMA, K = 32 bytes; issue-only / issue+drain over 512 MMAs)\n");
    printf("%-6s %-4s %-4s %-5s %-5s %-6s %10s %10s\n", "kind", "M", "N", "mode", "accs", "grid", "issue", "total");
    const int Ms[2] = {128, 64};
    const int Ns[7] = {8, 16, 32, 64, 128, 192, 256};
    for (int kind = 0; kind < 3; ++kind)
        for (int mi = 0; mi < 2; ++mi)
            for (int ni = 0; ni < 7; ++ni)
                for (int mode = 0; mode < 2; ++mode)
                    for (int grid : {1, 148}) {
                        const int M = Ms[mi], N = Ns[ni];
                        if (N == 8 && M == 128) continue;   // M=128 needs N % 16 == 0
                        if (kind == 2 && mode == 1) continue;
                        if (kind != 0 && (grid != 1 || M == 64)) continue;
                        if (mode == 1 && (M == 64 || grid != 1)) continue;
                        for (int accs : {1, 2}) {
                            if (accs == 2 && (N > 256 || mode == 1 || grid != 1)) continue;
                            const uint32_t id = idesc_for(kind, M, N);
                            long long h[2];
                            const int n_mma = 512;
                            if (kind == 0) mma_cost_kernel<KIND_I8><<<grid, 128, smem>>>(id, n_mma, mode, accs, d_out);
                            else if (kind == 1) mma_cost_kernel<KIND_F16><<<grid, 128, smem>>>(id, n_mma, mode, accs, d_out);
                            else mma_cost_kernel<KIND_TF32><<<grid, 128, smem>>>(id, n_mma, mode, accs, d_out);
                            cudaError_t e = cudaDeviceSynchronize();
                            if (e != cudaSuccess) { printf("  kind %d M %d N %d mode %d: %s\n", kind, M, N, mode, cudaGetErrorString(e)); exit(1); }
                            CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
                            printf("%-6s %-4d %-4d %-5s %-5d %-6d %10.1f %10.1f\n", kind == 0 ? "i8" : (kind == 1 ? "f16" : "tf32"), M, N,
                                   mode ? "TS" : "SS", accs, grid, h[0] / double(n_mma), h[1] / double(n_mma));
                        }
                    }
    printf("== 1b. same, A descriptor starting `shift` rows (x 128 B) into the tile (i8, M=128, SS)\n");
    for (int N : {32, 64, 128, 256})
        for (int shift : {0, 1, 2, 3, 4, 8, 9, 16, 17, 18}) {
            long long h[2];
            mma_cost_kernel<KIND_I8><<<1, 128, smem>>>(idesc_for(0, 128, N), 512, 0, 1, d_out, shift);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
            printf("N %-4d shift %-3d : %8.1f clk/MMA\n", N, shift, h[1] / 512.0);
        }
    cudaFree(d_out);
}

// ------------------------------------------------------------------------------------------------ 2. TMA ingest
struct IngestParams {
    int iters, depth, mode;      // mode 0 tiled 2-D, 1 im2col 3x3 walk, 2 tiled multicast
    int box_rows;                // rows per load (128 / cluster size for multicast)
    int csz;
    int tiles_total;             // wrap
    int HoWo, Wo;                // im2col geometry
    int inner_bytes;             // bytes per row (128 | 64)
};

__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint32_t bar_sa, uint32_t dst_sa, int32_t c0, int32_t c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst_sa), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_sa), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) ingest_kernel(const __grid_constant__ CUtensorMap map, const IngestParams p, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[16];
    const uint32_t rank = p.csz > 1 ? cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.depth; ++i) mbar_init(&full[i], 1);
        fence_mbar_init();
        tma_prefetch_desc(&map);
    }
    __syncthreads();
    if (p.csz > 1) cluster_sync_all();
    const int tile_bytes = 128 * p.inner_bytes;
    if (threadIdx.x == 0) {
        const int cluster_id = blockIdx.x / p.csz;
        const uint32_t smem_sa = smem_u32(smem), full_sa = smem_u32(full);
        long long first = 0;
        auto issue = [&](int i) {
            const int st = i % p.depth;
            const int tile = (cluster_id * 977 + i) % p.tiles_total;
            mbar_arrive_expect_tx_sa(full_sa + 8 * st, tile_bytes);
            if (p.mode == 0) {
                tma_load_2d_sa(&map, full_sa + 8 * st, smem_sa + st * tile_bytes, 0, tile * 128);
            } else if (p.mode == 2) {
                tma_load_2d_mc(&map, full_sa + 8 * st, smem_sa + st * tile_bytes + rank * p.box_rows * p.inner_bytes, 0,
                               tile * 128 + rank * p.box_rows, static_cast<uint16_t>((1u << p.csz) - 1));
            } else {
                // 3x3 pad-1 walk: consecutive loads = the 9 taps of one 128-pixel tile
                const int t = (cluster_id * 977 + i / 9) % p.tiles_total, tap = i % 9;
                const int m0 = t * 128;
                const int n = m0 / p.HoWo, rem = m0 - n * p.HoWo, p0 = rem / p.Wo, q0 = rem - p0 * p.Wo;
                tma_load_im2col_4d_sa(&map, full_sa + 8 * st, smem_sa + st * tile_bytes, 0, q0 - 1, p0 - 1, n,
                                      static_cast<uint16_t>(tap % 3), static_cast<uint16_t>(tap / 3));
            }
        };
        const long long t0 = clock64();
        for (int i = 0; i < p.depth; ++i) issue(i);
        for (int i = 0; i < p.iters; ++i) {
            const int st = i % p.depth;
            mbar_wait_sa(full_sa + 8 * st, (i / p.depth) & 1);
            if (i == 0) first = clock64() - t0;
            if (i + p.depth < p.iters) issue(i + p.depth);
        }
        const long long t1 = clock64();
        out[2 * blockIdx.x] = t1 - t0;
        out[2 * blockIdx.x + 1] = first;
    }
    __syncthreads();
    if (p.csz > 1) cluster_sync_all();
}

static void run_ingest() {
    printf("== 2. TMA ingest per SM from L2 (bytes/clk per CTA: median over CTAs; first = first-load latency in clk)\n");
    const size_t rows = 262144;   // x 128 B = 32 MiB (L2 resident)
    uint8_t* buf;
    CK(cudaMalloc(&buf, rows * 128));
    CK(cudaMemset(buf, 1, rows * 128));
    long long* d_out;
    CK(cudaMalloc(&d_out, 2 * 148 * 8 * sizeof(long long)));
    CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    printf("%-10s %-6s %-6s %-5s %-5s %10s %10s %10s\n", "mode", "inner", "depth", "csz", "grid", "B/clk med", "B/clk min", "first");
    auto run = [&](int mode, int inner, int depth, int csz, int grid) {
        CUtensorMap map;
        IngestParams p{};
        p.iters = 360; p.depth = depth; p.mode = mode; p.csz = csz; p.inner_bytes = inner;
        p.box_rows = 128 / csz;
        if (mode == 1) {
            // NHWC [N][56][56][inner], N so that it fills ~24 MiB
            const int H = 56, W = 56, N = static_cast<int>(rows * 128 / (size_t(H) * W * inner) * 3 / 4);
            cuuint64_t dims[4] = {cuuint64_t(inner), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
            cuuint64_t strides[3] = {cuuint64_t(inner), cuuint64_t(W) * inner, cuuint64_t(H) * W * inner};
            int lower[2] = {-1, -1}, upper[2] = {-1, -1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = g_im2col(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, buf, dims, strides, lower, upper, inner, 128, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("im2col encode failed %d\n", int(r)); return; }
            p.HoWo = H * W; p.Wo = W; p.tiles_total = N * H * W / 128;
        } else {
            const size_t r_total = rows * 128 / inner;
            cuuint64_t dims[2] = {cuuint64_t(inner), cuuint64_t(r_total)};
            cuuint64_t strides[1] = {cuuint64_t(inner)};
            cuuint32_t box[2] = {cuuint32_t(inner), cuuint32_t(p.box_rows)};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = g_tiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("tiled encode failed %d\n", int(r)); return; }
            p.tiles_total = static_cast<int>(r_total / 128);
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = depth * 128 * inner + 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) {   // first pass warms L2
            CK(cudaLaunchKernelEx(&cfg, ingest_kernel, map, p, d_out));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("ingest mode %d failed: %s\n", mode, cudaGetErrorString(e)); exit(1); }
        }
        std::vector<long long> h(2 * grid);
        CK(cudaMemcpy(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost));
        std::vector<double> bpc(grid);
        double fsum = 0;
        for (int i = 0; i < grid; ++i) { bpc[i] = double(p.iters) * 128 * inner / double(h[2 * i]); fsum += h[2 * i + 1]; }
        std::sort(bpc.begin(), bpc.end());
        printf("%-10s %-6d %-6d %-5d %-5d %10.1f %10.1f %10.0f\n", mode == 0 ? "tiled" : (mode == 1 ? "im2col3x3" : "multicast"), inner,
               depth, csz, grid, bpc[grid / 2], bpc[0], fsum / grid);
    };
    for (int grid : {1, 104, 148}) {
        for (int depth : {2, 4, 6, 12}) run(0, 128, depth, 1, grid);
        for (int depth : {4, 6, 12}) run(1, 128, depth, 1, grid);
        run(0, 64, 12, 1, grid);
        run(1, 64, 12, 1, grid);
    }
    for (int csz : {2, 4, 8}) {
        for (int depth : {4, 12}) {
            run(2, 128, depth, csz, csz);
            run(2, 128, depth, csz, 104 / csz * csz);
            run(2, 128, depth, csz, 144 / csz * csz);
        }
    }
    cudaFree(buf);
    cudaFree(d_out);
}

// ------------------------------------------------------------------------------------------------ 3. row-shifted descriptor
// A slab: 160 rows x 128 B (SW128, written with the absolute-address swizzle a TMA load would use). For a shift s the MMA reads
// rows [s, s+128). B: 32 rows x 128 B. D[128 x 32] = sum_k A[s+i][k] * B[j][k], checked against the host.
__global__ void __launch_bounds__(128, 1) shift_kernel(const int8_t* a_rows, const int8_t* b_rows, int shift, int base_off_mode, int32_t* d_out, int rb) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    uint8_t* a_s = smem;              // 160 x 128 = 20480 B
    uint8_t* b_s = smem + 20480;      // 32 x 128
    const int cpr = rb >> 4;                                  // 16-byte chunks per row
    const int lg = rb == 128 ? 7 : (rb == 64 ? 6 : 5);
    for (int i = threadIdx.x; i < 160 * cpr; i += blockDim.x) {
        const int row = i / cpr, c16 = i % cpr, sw = (row >> (7 - lg)) & (cpr - 1);
        *reinterpret_cast<uint4*>(a_s + row * rb + ((c16 ^ sw) << 4)) = *reinterpret_cast<const uint4*>(a_rows + row * 128 + c16 * 16);
    }
    for (int i = threadIdx.x; i < 32 * cpr; i += blockDim.x) {
        const int row = i / cpr, c16 = i % cpr, sw = (row >> (7 - lg)) & (cpr - 1);
        *reinterpret_cast<uint4*>(b_s + row * rb + ((c16 ^ sw) << 4)) = *reinterpret_cast<const uint4*>(b_rows + row * 128 + c16 * 16);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<32>(&tmem_ptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(2, 1, 1, 128, 32);
        const uint32_t bo = base_off_mode ? static_cast<uint32_t>(shift & 7) : 0u;
        const uint32_t lt = rb == 128 ? 2u : (rb == 64 ? 4u : 6u);
        const uint32_t hi_a = ((8u * rb) >> 4) | (1u << 14) | (bo << 17) | (lt << 29);
        const uint32_t hi_b = ((8u * rb) >> 4) | (1u << 14) | (lt << 29);
        const uint32_t a16 = (smem_u32(a_s) + shift * rb) >> 4, b16 = smem_u32(b_s) >> 4;
        for (int q = 0; q < rb / 32; ++q)
            tc_mma_lohi<KIND_I8>(tmem, ((a16 + 2 * q) & 0x3FFF) | (1u << 16), hi_a, ((b16 + 2 * q) & 0x3FFF) | (1u << 16), hi_b, idesc, q > 0);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[16];
    for (int c0 = 0; c0 < 32; c0 += 16) {
        tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) d_out[(warp * 32 + lane) * 32 + c0 + i] = static_cast<int32_t>(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<32>(tmem); }
}

static void run_shift() {
    printf("== 3. row-shifted SW128 descriptor (A start = tile + shift*128 B): mismatches out of 4096\n");
    std::vector<int8_t> a(160 * 128), b(32 * 128);
    srand(7);
    for (auto& x : a) x = static_cast<int8_t>(rand() % 255 - 127);
    for (auto& x : b) x = static_cast<int8_t>(rand() % 255 - 127);
    int8_t *da, *db;
    int32_t* dd;
    CK(cudaMalloc(&da, a.size())); CK(cudaMalloc(&db, b.size())); CK(cudaMalloc(&dd, 128 * 32 * 4));
    CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
    for (int rb : {128, 64, 32})
    for (int mode = 0; mode < 2; ++mode)
        for (int shift : {0, 1, 2, 3, 7, 8, 9, 16, 18, 31}) {
            if (mode == 1 && rb != 128) continue;
            CK(cudaMemset(dd, 0xff, 128 * 32 * 4));
            shift_kernel<<<1, 128, 20480 + 4096 + 1024>>>(da, db, shift, mode, dd, rb);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
            std::vector<int32_t> h(128 * 32);
            CK(cudaMemcpy(h.data(), dd, h.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 32; ++j) {
                    int32_t acc = 0;
                    for (int k = 0; k < rb; ++k) acc += int32_t(a[(i + shift) * 128 + k]) * int32_t(b[j * 128 + k]);
                    bad += acc != h[i * 32 + j];
                }
            printf("  row %3d B  base_offset %-8s shift %-3d : %d\n", rb, mode ? "shift&7" : "0", shift, bad);
        }
    cudaFree(da); cudaFree(db); cudaFree(dd);
}

int main(int argc, char** argv) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
    g_im2col = reinterpret_cast<PFN_encodeIm2col>(fn);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const char* which = argc > 1 ? argv[1] : "all";
    if (!strcmp(which, "all") || !strcmp(which, "shift")) run_shift();
    if (!strcmp(which, "all") || !strcmp(which, "mma")) run_mma_cost();
    if (!strcmp(which, "all") || !strcmp(which, "ingest")) run_ingest();
    return 0;
}
