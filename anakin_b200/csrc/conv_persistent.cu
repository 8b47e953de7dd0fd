// Persistent, tile-pipelined variant of the TMA-im2col implicit-GEMM convolution (conv_igemm.cu) for grids of several
// waves (large batches, the 112 x 112 stem, the 56 x 56 stage): one CTA per SM walks the tile list
//     tile t = (n-tile t / tiles_m, m-tile t % tiles_m),  t = blockIdx.x, blockIdx.x + gridDim.x, ...
// with the three roles running as free pipelines over the whole list instead of per tile:
//   * warp 0 (one lane): TMA producer -- the operand ring never drains between tiles, so no tile pays the ~1 us
//     first-load latency again and the prologue (barrier init, TMEM allocation, descriptor prefetch) is paid once;
//   * warp 1 (one lane): MMA issuer -- TWO accumulators in TMEM (2 x BN columns); tile i+1's MMAs run into one while
//   * warps 2..9: the epilogue reads tile i out of the other (tcgen05.ld -> bias / scale / residual / relu /
//     requantise -> swizzled staging tile in its OWN shared memory -> TMA store), so the epilogue -- the longest phase
//     of the wide 1x1 layers -- is hidden behind the next tile's main loop.
// Same arithmetic, same epilogue code (conv_common.cuh) and same tensor maps as the per-tile kernel: results are
// bit-identical. Reference loop being replaced: one SASS / cuDNN launch per layer with one CTA per tile
// (saber/funcs/impl/cuda/saber_conv.cpp:17-585, sass_funcs.h:481-555).
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "conv_common.cuh"

namespace b200 {

// smem: [ring: stages x (A 16 KiB + B BN*128 B)][staging 128 x BN x out_es][residual 128 x BN x res_es]
//       [bias | scale][full[MAX] empty[MAX] tmem_full[2] tmem_empty[2] res_full res_empty][tmem ptr]
__host__ __device__ constexpr int persistent_tail_bytes(int bn) { return 2 * bn * 4 + (2 * MAX_STAGES + 6) * 8 + 16; }

template <int KIND, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                       const ConvKParams p, const uint32_t idesc, const int tiles_m, const int tiles_total) {
    constexpr int SB = stage_bytes(BN, false);
    constexpr int B_OFF = A_STAGE_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;     // two accumulators (power of two: BN is)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* out_tile = smem + p.stages * SB;
    uint8_t* res_tile = out_tile + BLOCK_M * BN * p.out_es;
    float* bias_s = reinterpret_cast<float*>(res_tile + p.res_panels * BLOCK_M * p.res_pw);
    float* scale_s = bias_s + BN;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(scale_s + BN);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;    // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
    uint64_t* res_full_bar = tmem_empty_bar + 2;
    uint64_t* res_empty_bar = res_full_bar + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_empty_bar + 1);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int subs_per_stage = STAGE_K_BYTES / p.chunk;
    const int num_stage_iters = (p.KS + subs_per_stage - 1) / subs_per_stage;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
        if (p.res_panels > 0) tma_prefetch_desc(&map_res);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], EPI_THREADS);
        }
        mbar_init(res_full_bar, 1);
        mbar_init(res_empty_bar, EPI_THREADS);
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    pdl_launch_dependents();

    if (warp_idx == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            const uint32_t b_sub_bytes = BN * p.chunk;
            const uint32_t a_sub_bytes = BLOCK_M * p.chunk;
            const uint32_t tx_per_sub = a_sub_bytes + b_sub_bytes;
            const uint32_t ring_sa = smem_u32(smem), full_sa0 = smem_u32(full_bar), empty_sa0 = smem_u32(empty_bar);
            const bool may_pad = p.KS != p.KS_real;
            const int res_cols_per_panel = p.res_panels > 0 ? p.res_pw / p.res_es : 0;
            int stage = 0;
            uint32_t phase = 0, stage_sa = ring_sa, full_sa = full_sa0, empty_sa = empty_sa0;
            uint32_t res_phase = 1;       // parity to wait on res_empty: the first use finds the buffer free
            pdl_wait_prior_grid();
            for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
                const int nt = tile / tiles_m, mt = tile - nt * tiles_m;
                const int m0 = mt * BLOCK_M, n0 = nt * BN;
                const int n_img = m0 / p.HoWo;
                const int rem = m0 - n_img * p.HoWo;
                const int p0 = rem / p.Wo;
                const int q0 = rem - p0 * p.Wo;
                const int base_w = q0 * p.stride_w - p.pad_w;
                const int base_h = p0 * p.stride_h - p.pad_h;
                int ks = 0, cc = 0, r = 0, s = 0;
                int c_coord = 0, k_coord = 0, off_w = 0, off_h = 0;
                for (int it = 0; it < num_stage_iters; ++it) {
                    const int nsub = min(subs_per_stage, p.KS - ks);
                    mbar_wait_sa(empty_sa, phase ^ 1);
                    mbar_arrive_expect_tx_sa(full_sa, nsub * tx_per_sub);
                    uint32_t a_dst = stage_sa, b_dst = stage_sa + B_OFF;
#pragma unroll 1
                    for (int j = 0; j < nsub; ++j) {
                        const bool pad_step = may_pad && ks >= p.KS_real;   // re-reads tap (0,0); its weights are zero
                        tma_load_im2col_4d_sa(&map_a, full_sa, a_dst, pad_step ? 0 : c_coord, base_w, base_h, n_img,
                                              static_cast<uint16_t>(pad_step ? 0 : off_w),
                                              static_cast<uint16_t>(pad_step ? 0 : off_h));
                        tma_load_2d_sa(&map_b, full_sa, b_dst, k_coord, n0);
                        a_dst += a_sub_bytes; b_dst += b_sub_bytes;
                        ++ks;
                        k_coord += p.chunk_el;
                        c_coord += p.chunk_el;
                        if (++cc == p.CC) {
                            cc = 0; c_coord = 0;
                            off_w += p.dil_w;
                            if (++s == p.S) { s = 0; off_w = 0; ++r; off_h += p.dil_h; }
                        }
                    }
                    stage_sa += SB; full_sa += 8; empty_sa += 8;
                    if (++stage == p.stages) { stage = 0; phase ^= 1; stage_sa = ring_sa; full_sa = full_sa0; empty_sa = empty_sa0; }
                    if (it == 0 && p.res_panels > 0) {
                        // the residual tile of this output tile: into the single residual buffer once the epilogue of
                        // the previous tile has read it out
                        mbar_wait(res_empty_bar, res_phase);
                        res_phase ^= 1;
                        mbar_arrive_expect_tx(res_full_bar, p.res_panels * BLOCK_M * p.res_pw);
                        for (int j = 0; j < p.res_panels; ++j)
                            tma_load_2d(&map_res, res_full_bar, res_tile + j * BLOCK_M * p.res_pw, n0 + j * res_cols_per_panel, m0);
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t lt = layout_type_for_chunk(p.chunk);
            const uint32_t a_sub16 = (BLOCK_M * p.chunk) >> 4, b_sub16 = (BN * p.chunk) >> 4;
            const bool swz = p.chunk >= 32;
            const uint32_t hi = (swz ? (8u * p.chunk) >> 4 : 128u >> 4) | (1u << 14) | (lt << 29);
            const uint32_t a_lbo = (swz ? 1u : a_sub16) << 16, b_lbo = (swz ? 1u : b_sub16) << 16;
            const int mma_per_sub = swz ? (p.chunk >> 5) : 1;
            const int sub_step = swz ? 1 : 2;
            const uint32_t ring16 = smem_u32(smem) >> 4;
            const uint32_t full_sa0 = smem_u32(full_bar), empty_sa0 = smem_u32(empty_bar);
            int stage = 0;
            uint32_t phase = 0, stage16 = ring16, full_sa = full_sa0, empty_sa = empty_sa0;
            uint32_t acc = 0, acc_phase[2] = {1, 1};      // parity to wait on tmem_empty: free at first use
            for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase[acc]);     // the epilogue has drained this accumulator
                acc_phase[acc] ^= 1;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accum = 0;
                int ks = 0;
                for (int it = 0; it < num_stage_iters; ++it) {
                    mbar_wait_sa(full_sa, phase);
                    tc_fence_after();
                    const int nsub = min(subs_per_stage, p.KS - ks);
                    uint32_t a16 = stage16, b16 = stage16 + (B_OFF >> 4);
                    for (int j = 0; j < nsub; j += sub_step) {
                        for (int q = 0; q < mma_per_sub; ++q) {
                            tc_mma_lohi<KIND>(d_tmem, (a16 + 2 * q) | a_lbo, hi, (b16 + 2 * q) | b_lbo, hi, idesc, accum);
                            accum = 1;
                        }
                        a16 += sub_step * a_sub16;
                        b16 += sub_step * b_sub16;
                    }
                    ks += nsub;
                    tc_commit_sa(empty_sa);
                    stage16 += SB >> 4; full_sa += 8; empty_sa += 8;
                    if (++stage == p.stages) { stage = 0; phase ^= 1; stage16 = ring16; full_sa = full_sa0; empty_sa = empty_sa0; }
                }
                tc_commit(&tmem_full_bar[acc]);
                acc ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp_idx & 3;
        const int row = quarter * 32 + lane;
        constexpr int COLS_PER_WARP = BN / (EPI_WARPS / 4);
        auto lg2 = [](int pw) { return pw == 128 ? 7 : (pw == 64 ? 6 : (pw == 32 ? 5 : 4)); };
        const PanelRow out_row = make_panel_row(smem_u32(out_tile), lg2(p.out_pw), row);
        const PanelRow res_row = make_panel_row(smem_u32(res_tile), lg2(p.res_pw ? p.res_pw : 128), row);
        const uint32_t bias_sa = smem_u32(bias_s), scale_sa = smem_u32(scale_s);
        const int cbeg = ((warp_idx - 2) >> 2) * COLS_PER_WARP, cend = cbeg + COLS_PER_WARP;
        const bool storer = warp_idx == 2 && lane == 0;
        const int cols_per_panel = p.out_pw / p.out_es;
        uint32_t acc = 0, full_phase[2] = {0, 0}, res_phase = 0;
        int n0_loaded = -1;
        for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
            const int nt = tile / tiles_m, mt = tile - nt * tiles_m;
            const int m0 = mt * BLOCK_M, n0 = nt * BN;
            // the staging tile is free once the previous tile's TMA store has read it; the tables follow the n-tile
            if (storer) tma_store_wait_read();
            if (n0 != n0_loaded) {
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // nobody still reads the old tables
                for (int i = threadIdx.x - 64; i < BN; i += EPI_THREADS) {
                    const bool ok = (n0 + i) < p.K;
                    bias_s[i] = (p.bias != nullptr && ok) ? __ldg(p.bias + n0 + i) : 0.f;
                    scale_s[i] = (p.scale != nullptr && ok) ? __ldg(p.scale + n0 + i) : 1.f;
                }
                n0_loaded = n0;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            mbar_wait(&tmem_full_bar[acc], full_phase[acc]);
            full_phase[acc] ^= 1;
            if (p.res_panels > 0) { mbar_wait(res_full_bar, res_phase); res_phase ^= 1; }
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cend; c0 += 16) {
                if (n0 + c0 >= p.K) break;
                uint32_t v0[16];
                tmem_ld_32x32b_x16(t_row + c0, v0);
                tmem_ld_wait();
                epilogue16<KIND>(p, v0, c0, bias_sa, scale_sa, res_row, out_row);
            }
            // accumulator and residual buffer are read out: hand them back before the store
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);
            if (p.res_panels > 0) mbar_arrive(res_empty_bar);
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            if (storer) {
                for (int j = 0; j < p.out_panels; ++j) {
                    if (n0 + j * cols_per_panel >= p.K) break;
                    tma_store_2d(&map_out, out_tile + j * BLOCK_M * p.out_pw, n0 + j * cols_per_panel, m0);
                }
                tma_store_commit();
            }
            acc ^= 1;
        }
        if (storer) tma_store_wait_read();
    }

    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int KIND, int BN>
static void launch_persistent(b200_conv_plan* pl, void* stream) {
    auto kern = conv_persistent_kernel<KIND, BN>;
    static std::atomic<bool> opted_in[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < kMaxDevices && !opted_in[dev].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
        opted_in[dev].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pl->persistent_ctas);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, pl->map_a, pl->map_b, pl->map_out, pl->map_res, pl->kp, pl->idesc,
                       static_cast<int>(pl->grid.x), static_cast<int>(pl->grid.x * pl->grid.y));
    count_launch();
}

template <int KIND>
static bool select_persistent(b200_conv_plan* pl) {
    switch (pl->bn) {
        case 32: pl->launch = launch_persistent<KIND, 32>; return true;
        case 64: pl->launch = launch_persistent<KIND, 64>; return true;
        case 128: pl->launch = launch_persistent<KIND, 128>; return true;
        case 256: pl->launch = launch_persistent<KIND, 256>; return true;
    }
    return false;
}

// Turn a finished per-tile im2col plan into the persistent variant when its grid spans several waves. Keeps BN, the
// tensor maps and the epilogue parameters; recomputes the ring depth for one CTA per SM with a dedicated staging tile.
bool persistent_plan_setup(b200_conv_plan* pl) {
    const b200_conv_desc_t& d = pl->desc;
    const char* env = getenv("B200_SABER_PERSISTENT");
    if (env && env[0] == '0') return false;
    if (d.math == B200_MATH_TF32X3 || pl->kp.split != 1) return false;
    const int sms = sm_count();
    const int tiles = static_cast<int>(pl->grid.x * pl->grid.y);
    const bool force = env && env[0] == '2';
    // worth it once every SM gets more than the two tiles that co-resident CTAs already overlap ...
    if (!force && tiles <= 2 * sms) return false;
    const int bn = pl->bn;
    {
        // ... and the tile is not epilogue-dominated: one persistent CTA has ONE set of 8 epilogue warps, two co-resident
        // per-tile CTAs have two, and for the short-K / wide-N layers (1x1 64 -> 256 + residual) the read-out is the
        // tile (measured, tools/ab_layers.py: 22.3 us per-tile vs 24.8 us persistent at batch 32)
        const double k_bytes = static_cast<double>(pl->g.KS) * pl->g.chunk;
        const double ingest = (BLOCK_M + bn) * k_bytes / 38.7;
        const double mma = k_bytes / 32.0 * (bn / 2.0 > 32.0 + bn / 4.0 ? bn / 2.0 : 32.0 + bn / 4.0);
        const double loop = mma > ingest ? mma : ingest;
        const double epi = bn * (pl->kp.res_es ? 11.0 : 9.0);
        // With many tiles per SM (>= 6) the per-tile CTAs pay their prologue / TMEM allocation / drain once per tile and
        // the walker wins further into epilogue-bound territory: MobileNet-v1 FP16 b16 conv2_sep (1x1 32 -> 64 on 200704
        // pixels, 1568 tiles, epi / loop = 1.8) 20.8 -> 14.2 us in-net; at 5.3 tiles per SM (ResNet-50 b32 res2a_branch2a,
        // same ratio) the step was 0.8 % slower, so the relaxed bound applies from 6 tiles per SM only.
        static const double epi_env = [] { const char* e = getenv("B200_SABER_PERSISTENT_EPI"); const double v = e ? atof(e) : 0.0; return v > 0.0 ? v : 0.0; }();
        const double epi_ratio = epi_env > 0.0 ? epi_env : (tiles >= 6 * sms ? 2.5 : 1.5);
        if (!force && epi > epi_ratio * loop) return false;
    }
    const int sb = stage_bytes(bn, false);
    const int staging = BLOCK_M * bn * pl->kp.out_es;
    const int res_bytes = BLOCK_M * bn * pl->kp.res_es;
    const int fixed = staging + res_bytes + persistent_tail_bytes(bn) + 1024;
    int stages = (MAX_SMEM - fixed) / sb;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return false;
    bool ok = false;
    if (d.math == B200_MATH_I8) ok = select_persistent<KIND_I8>(pl);
    else if (d.math == B200_MATH_F16) ok = select_persistent<KIND_F16>(pl);
    else ok = select_persistent<KIND_TF32>(pl);
    if (!ok) return false;
    pl->kp.stages = stages;
    pl->smem_bytes = stages * sb + fixed;
    pl->persistent_ctas = tiles < sms ? tiles : sms;
    pl->persistent = true;
    return true;
}

}  // namespace b200
