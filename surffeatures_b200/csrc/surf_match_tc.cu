// surf_match_tc.cu -- stage 5 on the 5th-generation tensor cores: candidate search for
// BFMatcher(NORM_L2)::knnMatch (SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:640).
//
// The one dense contraction of the path is S = Q * T^T.  It runs as tcgen05.mma (kind::f16, BF16
// operands, FP32 accumulators in TMEM) with a hi/lo BF16 split of the fp32 descriptors
// (q.t ~ qh.th + qh.tl + ql.th), so the approximate squared distance |t|^2 - 2 q.t is good to ~1e-5.
// It is used ONLY to pick kCand candidates per query; the reported distances and the final order
// come from exact fp32 rescoring of those candidates (same arithmetic as the exact kernel in
// surf_match.cu), and a per-query margin test proves the candidates contain the true top-k;
// queries that fail the test are redone by the exact kernel, so results equal the exact path.
//
//   k_tc_prep        fp32 rows -> BF16 hi / lo in the UMMA canonical K-major layout + |row|^2
//   k_tc_candidates    128 x 256 output tiles: operands copied to shared memory, one elected thread
//                      issues tcgen05.mma, accumulators read back with tcgen05.ld, per-row running
//                      top-kCand kept in registers across the train tiles (filtered lists: k = 3, 4, databases)
//   k_tc_candidates3w  k <= 2 against a frame: warp-specialised (MMA warp, copy warp, 16 epilogue warps), four
//                      operand stages and four TMEM accumulators, scores as sortable 32-bit keys, branch-free lists
//   k_tc_candidates3   the same key epilogue with MMA and epilogue back to back (128-d rows, A/B)
//   k_tc_rescore       exact fp32 distances of the candidates, top-k, margin test
//   k_tc_redo          exact scan for the queries the margin test could not prove
#include <cuda_bf16.h>

#include <cfloat>
#include <cstdlib>

#include "surf_internal.h"
#include "surf_ptx.cuh"

namespace surfb200 {

constexpr int kTM = 128;  // queries per tile  (UMMA M)
constexpr int kTN = 256;  // train rows per tile (UMMA N)
constexpr uint32_t kTmemCols = 256;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
// Bounded wait: a barrier that never completes traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (!mbar_wait_bounded(bar, parity)) __trap();
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, BF16 x BF16 -> FP32, issued by one thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// the same with the accumulate flag fixed at compile time (no predicate computed from a register)
__device__ __forceinline__ void umma_bf16_first(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, K-major, no swizzle (canonical layout ((8,m),(8,2k)) with 16-byte
// rows inside an 8 x 16 B core matrix): LBO = byte step between the two core matrices along K,
// SBO = byte step between 8-row groups.  Bits 46-47 = 1 is the sm_100 descriptor version.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// Instruction descriptor for kind::f16: D = F32, A = B = BF16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// K5a: prep.  Row r, element k of a slot lives at bf16 index ((r/8)*(D/8) + k/8)*64 + (r%8)*8 + k%8,
// so any 8-aligned block of rows is one contiguous run of core matrices.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_tc_prep(const TcSegment *__restrict__ segs, int n_segs, TcPool pool)
{
    constexpr int G = D / 8;  // 8-element groups per row
    const int seg = blockIdx.y;
    if (seg >= n_segs) return;
    const TcSegment sg = segs[seg];
    const float *src = sg.src;
    const int base = sg.dst_row0;  // live rows already in the slot (append); 0 for a fresh slot
    const int n = min(sg.off ? sg.off[1] - sg.off[0] : sg.n, pool.slot_rows - base);
    const int src_row0 = sg.off ? sg.off[0] : sg.row0;
    // rows base + [0, npad) are written; base + npad is the next multiple of 256 (padding rows get |t|^2 = inf)
    const int npad = min((base + n + kTN - 1) / kTN * kTN, pool.slot_rows) - base;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = t / G, g = t - r * G;
    if (t == 0) pool.count[sg.slot] = base + n;
    if (r >= npad) return;  // G divides the warp: the lanes of a row leave together
    float v[8];
    if (r < n) {
        const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)(src_row0 + r) * D + g * 8);
        const float4 a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = 0.f;
    }
    uint32_t hi[4], lo[4];
    float nrm = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[k]), h1 = __float2bfloat16_rn(v[k + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[k] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[k + 1] - __bfloat162float(h1));
        hi[k / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lo[k / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        nrm = fmaf(v[k], v[k], nrm);
        nrm = fmaf(v[k + 1], v[k + 1], nrm);
    }
    // the G lanes of a row are consecutive lanes of one warp (G = 8 or 16)
    const unsigned gm = (G == 32 ? 0xffffffffu : ((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) nrm += __shfl_xor_sync(gm, nrm, d);
    const size_t slot_base = (size_t)sg.slot * pool.slot_rows;
    const int R = base + r;
    const size_t e = (slot_base + (size_t)(R & ~7)) * D + (size_t)g * 64 + (size_t)(R & 7) * 8;
    *reinterpret_cast<uint4 *>(pool.hi + e) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(pool.lo + e) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (g == 0) {
        pool.norm[slot_base + R] = r < n ? nrm : INFINITY;
        if (r < n) atomicMax(reinterpret_cast<int *>(pool.max_norm + sg.slot), __float_as_int(nrm));
    }
}

__global__ void k_tc_reset(TcPool pool, const TcSegment *__restrict__ segs, int n_segs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_segs) pool.max_norm[segs[i].slot] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// K5b: candidates.
// ------------------------------------------------------------------------------------------------
// Running candidate set of one query row: KC entries, unsorted; `worst` is the largest kept
// approximate value and `wslot` where it sits.  A new value replaces the worst (strictly smaller
// only, so equal values keep the earlier train row).  All indexing is by predicated selects so the
// arrays stay in registers.  KC = 4 serves k <= 2 (the candidate scan is instruction-bound: every
// replacement costs the whole warp; KC = 3 was measured and sends ~0.7 % of the queries to the redo
// kernel because the bound is then the third neighbour itself), KC = 8 serves k = 3, 4.
template <int KC>
struct Cand {
    float d[KC];
    int i[KC];
    float worst;
    int wslot;
};

template <int KC>
__device__ __forceinline__ void cand_replace_worst(Cand<KC> &c, float dist, int idx)
{
#pragma unroll
    for (int r = 0; r < KC; r++) {
        const bool hit = (r == c.wslot);
        c.d[r] = hit ? dist : c.d[r];
        c.i[r] = hit ? idx : c.i[r];
    }
    float w = c.d[0];
    int ws = 0;
#pragma unroll
    for (int r = 1; r < KC; r++) {
        const bool g = c.d[r] > w;
        w = g ? c.d[r] : w;
        ws = g ? r : ws;
    }
    c.worst = w;
    c.wslot = ws;
}

constexpr int kCandThreads = 512;  // 16 warps: warp w owns TMEM lanes 32*(w%4).. and the column quarter w/4

// The filtered-list candidate kernel (any k <= 4, any number of train rows): one 256-column accumulator per CTA,
// MMA(t) -> epilogue(t) back to back, two CTAs per SM overlap each other.  It serves k = 3, 4 and train sets of more
// than 2^14 rows per split (a database); for k <= 2 against a frame's keypoints the key kernels below replace it.
template <int D, int KC>
__global__ void __launch_bounds__(kCandThreads, (D == 64 && KC <= 4) ? 2 : 1) k_tc_candidates(TcPool pool, TcPool tpool, const TcProblem *__restrict__ probs,
                                                       int n_probs, int mt_per_prob, int n_splits, int cand_stride,
                                                       TcCand *__restrict__ cand_out, int *__restrict__ work_head)
{
    constexpr int KSTEPS = D / 16;
    constexpr uint32_t SBO = (D / 8) * 128;        // bytes between 8-row groups
    constexpr uint32_t A_BYTES = kTM * D * 2;      // one of hi / lo
    constexpr uint32_t B_BYTES = kTN * D * 2;
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *sA_hi = sm, *sA_lo = sm + A_BYTES;
    unsigned char *sB_hi = sm + 2 * A_BYTES, *sB_lo = sB_hi + B_BYTES;
    float *sTn = reinterpret_cast<float *>(sB_lo + B_BYTES);  // |t|^2 of the train tile, two stages
    __shared__ __align__(8) uint64_t s_bar[2];  // [0] MMA done (tcgen05.commit), [1] operands landed (bulk copies)
    __shared__ uint32_t s_tmem;
    __shared__ int s_item;
    __shared__ float s_thr[4][kTM];  // filter threshold published by each column quarter of a query row

    const int tid = threadIdx.x, warp = tid >> 5;
    const int quarter = warp >> 2, qrow = (warp & 3) * 32 + (tid & 31);
    volatile float *my_thr = &s_thr[quarter][qrow];
    const volatile float *row_thr = &s_thr[0][qrow];  // the four quarters' thresholds of this row, kTM apart
    const uint32_t bar_mma = smem_u32(&s_bar[0]), bar_ld = smem_u32(&s_bar[1]);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_ld, 1);
        mbar_init_fence();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc(kTM, kTN);
    uint32_t ph_mma = 0, ph_ld = 0;
    const int total_items = n_probs * mt_per_prob * n_splits;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(work_head, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total_items) break;
        const int p = item / (mt_per_prob * n_splits);
        const int rem = item - p * (mt_per_prob * n_splits);
        const int mt = rem / n_splits, split = rem - mt * n_splits;
        const TcProblem pr = probs[p];
        const int nq = pool.count[pr.q_slot], nt = tpool.count[pr.t_slot];
        const int q0 = mt * kTM;
        if (q0 >= nq) continue;
        // train tiles of this split
        const int n_tiles = (nt + kTN - 1) / kTN;
        const int per = (n_tiles + n_splits - 1) / n_splits;
        const int tile0 = split * per, tile1 = min(n_tiles, tile0 + per);

        Cand<KC> best;
#pragma unroll
        for (int r = 0; r < KC; r++) {
            best.d[r] = INFINITY;
            best.i[r] = -1;
        }
        best.worst = INFINITY;
        best.wslot = 0;
        float thr = INFINITY;
        *my_thr = INFINITY;  // read by the partner only after the tile loop's first __syncthreads
        const size_t trow_base = (size_t)pr.t_slot * tpool.slot_rows;
        // operands arrive by TMA bulk copies (contiguous runs of core matrices), counted on bar_ld
        auto load_train_tile = [&](int tile, int stage) {
            const size_t row0 = trow_base + (size_t)tile * kTN;
            bulk_g2s(smem_u32(sB_hi), tpool.hi + row0 * D, B_BYTES, bar_ld);
            bulk_g2s(smem_u32(sB_lo), tpool.lo + row0 * D, B_BYTES, bar_ld);
            bulk_g2s(smem_u32(sTn + stage * kTN), tpool.norm + row0, kTN * sizeof(float), bar_ld);
        };
        if (tid == 0 && tile0 < tile1) {
            const size_t e0 = ((size_t)pr.q_slot * pool.slot_rows + q0) * D;
            mbar_expect_tx(bar_ld, 2 * A_BYTES + 2 * B_BYTES + kTN * sizeof(float));
            bulk_g2s(smem_u32(sA_hi), pool.hi + e0, A_BYTES, bar_ld);
            bulk_g2s(smem_u32(sA_lo), pool.lo + e0, A_BYTES, bar_ld);
            load_train_tile(tile0, 0);
        }

        for (int tile = tile0; tile < tile1; tile++) {
            const int t0 = tile * kTN;
            const int stage = (tile - tile0) & 1;
            mbar_wait(bar_ld, ph_ld);  // operands of this tile are in shared memory
            ph_ld ^= 1;
            __syncthreads();  // everyone has observed this phase before the next one can be armed
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = smem_u32(sA_hi), a_lo = smem_u32(sA_lo);
                const uint32_t b_hi = smem_u32(sB_hi), b_lo = smem_u32(sB_lo);
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ks++) {
                    const uint32_t ko = ks * 256;  // two core matrices (16 K-elements) per step
                    umma_bf16(tmem, umma_desc(a_hi + ko, 128, SBO), umma_desc(b_hi + ko, 128, SBO), idesc, ks > 0);
                }
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ks++) {
                    const uint32_t ko = ks * 256;
                    umma_bf16(tmem, umma_desc(a_hi + ko, 128, SBO), umma_desc(b_lo + ko, 128, SBO), idesc, 1);
                    umma_bf16(tmem, umma_desc(a_lo + ko, 128, SBO), umma_desc(b_hi + ko, 128, SBO), idesc, 1);
                }
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma);  // accumulators complete; the train tile in shared memory is free
            ph_mma ^= 1;
            tc_fence_after();
            if (tid == 0 && tile + 1 < tile1) {  // prefetch the next train tile under this tile's epilogue
                mbar_expect_tx(bar_ld, 2 * B_BYTES + kTN * sizeof(float));
                load_train_tile(tile + 1, stage ^ 1);
            }

            // ---- epilogue: 16 warps; warp w reads TMEM lanes 32*(w%4).. (its query rows) and the column
            //      quarter w/4, so four threads share a query row and each keeps its own list.  The filter
            //      threshold `thr` is the smallest of the four lists' worst kept values: each of them is >= the
            //      KC-th smallest value of the whole row, so nothing that belongs to the row's KC smallest is
            //      ever rejected, and every value a list rejects or evicts is >= the final thr.
            const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            const float *tn = sTn + stage * kTN;
            const int lim = min(kTN, nt - t0);
            const int cbeg = quarter * (kTN / 4);
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + kTN / 4; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + c0, v);
                thr = fminf(fminf(thr, row_thr[0]), fminf(row_thr[kTM], fminf(row_thr[2 * kTM], row_thr[3 * kTM])));
                if (c0 < lim) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 t4 = *reinterpret_cast<const float4 *>(tn + c0 + c);
                        // |t|^2 - 2 q.t  (+ |q|^2 later); padded train rows carry |t|^2 = inf
                        const float d0 = fmaf(-2.f, v[c], t4.x), d1 = fmaf(-2.f, v[c + 1], t4.y);
                        const float d2 = fmaf(-2.f, v[c + 2], t4.z), d3 = fmaf(-2.f, v[c + 3], t4.w);
                        if (fminf(fminf(d0, d1), fminf(d2, d3)) < thr) {  // one test per four columns
                            const int idx = t0 + c0 + c;
                            if (d0 < thr) { cand_replace_worst(best, d0, idx); thr = fminf(thr, best.worst); }
                            if (d1 < thr) { cand_replace_worst(best, d1, idx + 1); thr = fminf(thr, best.worst); }
                            if (d2 < thr) { cand_replace_worst(best, d2, idx + 2); thr = fminf(thr, best.worst); }
                            if (d3 < thr) { cand_replace_worst(best, d3, idx + 3); thr = fminf(thr, best.worst); }
                        }
                    }
                }
                *my_thr = thr;
            }
            tc_fence_before();
            __syncthreads();  // every warp has drained TMEM before the next tile's MMA overwrites it
        }
        // ---- merge the four quarter lists of every query row into one list of the KC smallest values.  The
        //      scratch aliases the train-tile buffer: every MMA that read it has completed (bar_mma), and the
        //      next item's bulk copies are issued only after the proxy fence and the barrier at the loop top.
        struct Ent {
            float d;
            int i;
        };
        static_assert(sizeof(Ent) * kTM * 4 * KC <= B_BYTES, "merge scratch must fit the train-tile buffer");
        Ent *s_merge = reinterpret_cast<Ent *>(sB_hi);  // [row][quarter][KC]
#pragma unroll
        for (int r = 0; r < KC; r++) s_merge[(qrow * 4 + quarter) * KC + r] = Ent{best.d[r], best.i[r]};
        *my_thr = thr;
        __syncthreads();
        if (tid < kTM) {
            const int row = tid;
            Cand<KC> m;
#pragma unroll
            for (int r = 0; r < KC; r++) {
                m.d[r] = INFINITY;
                m.i[r] = -1;
            }
            m.worst = INFINITY;
            m.wslot = 0;
            for (int e = 0; e < 4 * KC; e++) {
                const Ent en = s_merge[row * 4 * KC + e];
                if (en.i >= 0 && en.d < m.worst) cand_replace_worst(m, en.d, en.i);
            }
            if (q0 + row < nq) {
                TcCand o;
#pragma unroll
                for (int r = 0; r < kCand; r++) {
                    o.d[r] = r < KC ? m.d[r < KC ? r : 0] : INFINITY;
                    o.i[r] = r < KC ? m.i[r < KC ? r : 0] : -1;
                }
                // lower bound of every value of this split that is not in the list: what the quarter lists
                // rejected or evicted (>= their final thresholds) and what the merge dropped (>= the list's worst)
                const float bound = fminf(fminf(s_thr[0][row], s_thr[1][row]), fminf(s_thr[2][row], s_thr[3][row]));
                o.worst = fminf(bound, m.worst);
                cand_out[(size_t)(p * n_splits + split) * cand_stride + q0 + row] = o;
            }
        }
        fence_async_smem();  // generic-proxy accesses to the tile buffer before the next bulk copy overwrites it
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}


size_t tc_candidates_smem(int dim) { return (size_t)(2 * kTM + 2 * kTN) * dim * 2 + 2 * kTN * sizeof(float) + 1024; }

// ------------------------------------------------------------------------------------------------
// K5b', the candidate kernel for k <= 2 and train splits of at most 2^14 rows (a frame's keypoints): the epilogue is
// BRANCH-FREE.  With 32 query rows in a warp, "some lane has a value under its row's threshold" is true for almost
// every group of columns when the train set is a frame (a few thousand rows), so the filtered list updates of
// k_tc_candidates run for the whole warp almost every time (12.6 instructions per score, measured).  Here a score is
// turned into ONE sortable 32-bit key and pushed through a sorted three-entry list with unsigned min / max only:
//   u    = sigma * (|t|^2 - 2 q.t + off)  in [0, 1)           sigma, off from the slots' largest norms
//   t'   = 2^(s-9) + u                    one FFMA: fma(acc, -2 sigma, tnB[col]), tnB = 2^(s-9) + sigma (|t|^2 + off)
//   key  = bits(t') << s  |  row of the train split (s bits)  one IMAD: the shift drops sign, exponent and the s - 9
//                                                             mantissa bits that are zero in [2^(s-9), 2^(s-9) + 1)
// i.e. the FFMA's own rounding quantises u to 32 - s bits (18 bits for 16 384 rows: 4e-6 of the value range) and the
// low s bits are free for the index.  The keys go into a sorted list (m1 <= m2 <= m3) either pair by pair (QUAD =
// false: two keys are sorted, then merged, 8 min / max per pair, three-input VIMNMX3 where two mins follow each other;
// m3 bounds everything the thread did not keep) or, the default, four at a time through a tournament
// (key_insert_quad).  Four threads share a query row (column quarters); at the end of an item their lists are merged:
// the four smallest keys are the candidates, min(fifth, the four threads' bounds) decoded and lowered by the
// quantisation slack is the bound the margin test of k_tc_rescore uses.  Padded train rows get the largest value of
// the range.  tests/test_match_keys_model.py restates the arithmetic and the bound in numpy and checks both.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void key_insert_pair(uint32_t &m1, uint32_t &m2, uint32_t &m3, uint32_t x, uint32_t y)
{
    const uint32_t lo = min(x, y), hi = max(x, y);
    const uint32_t a = max(m1, lo), b = max(m2, lo), c = max(m1, hi);
    m3 = min(m3, min(b, c));
    m2 = min(m2, min(a, hi));
    m1 = min(m1, lo);
}

// Four keys at a time: a tournament finds the smallest, which alone goes through the sorted list; the other three are
// only folded into `bnd`, a lower bound of everything the thread dropped without listing (their minimum is the quad's
// second smallest).  13 min / max per four scores instead of 16.  A true neighbour that loses its quad to a better one
// (3 of 4 352 columns share a quad with the best) is then not a candidate, the bound says so, and the query is redone.
__device__ __forceinline__ void key_insert_quad(uint32_t &m1, uint32_t &m2, uint32_t &m3, uint32_t &bnd, uint32_t k0, uint32_t k1,
                                                uint32_t k2, uint32_t k3)
{
    const uint32_t lo1 = min(k0, k1), hi1 = max(k0, k1), lo2 = min(k2, k3), hi2 = max(k2, k3);
    const uint32_t L = min(lo1, lo2), M = max(lo1, lo2);
    bnd = min(min(bnd, hi1), min(hi2, M));
    const uint32_t t1 = max(m1, L);
    m1 = min(m1, L);
    const uint32_t t2 = max(m2, t1);
    m2 = min(m2, t1);
    m3 = min(m3, t2);
}

size_t tc_candidates3_smem(int dim) { return (size_t)(2 * kTM + 2 * kTN) * dim * 2 + 4 * kTN * sizeof(float) + 1024; }

template <int D, bool QUAD>
__global__ void __launch_bounds__(kCandThreads, D == 64 ? 2 : 1) k_tc_candidates3(TcPool pool, TcPool tpool, const TcProblem *__restrict__ probs,
                                                        int n_probs, int mt_per_prob, int n_splits, int cand_stride,
                                                        TcCand *__restrict__ cand_out, int *__restrict__ work_head, int idx_bits, uint32_t key_mul)
{
    constexpr int KSTEPS = D / 16;
    constexpr uint32_t SBO = (D / 8) * 128;        // bytes between 8-row groups
    constexpr uint32_t A_BYTES = kTM * D * 2;      // one of hi / lo
    constexpr uint32_t B_BYTES = kTN * D * 2;
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *sA_hi = sm, *sA_lo = sm + A_BYTES;
    unsigned char *sB_hi = sm + 2 * A_BYTES, *sB_lo = sB_hi + B_BYTES;
    float *sTn = reinterpret_cast<float *>(sB_lo + B_BYTES);  // |t|^2 of the train tile, two stages (bulk copies)
    float *sTnB = sTn + 2 * kTN;                               // 2^(s-9) + sigma (|t|^2 + off), two stages (generic)
    __shared__ __align__(8) uint64_t s_bar[2];  // [0] MMA done (tcgen05.commit), [1] operands landed (bulk copies)
    __shared__ uint32_t s_tmem;
    __shared__ int s_item;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int quarter = warp >> 2, qrow = (warp & 3) * 32 + (tid & 31);
    const uint32_t bar_mma = smem_u32(&s_bar[0]), bar_ld = smem_u32(&s_bar[1]);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_ld, 1);
        mbar_init_fence();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc(kTM, kTN);
    uint32_t ph_mma = 0, ph_ld = 0;
    const int total_items = n_probs * mt_per_prob * n_splits;
    const int vbits = 32 - idx_bits;                                  // bits of the quantised value
    const uint32_t pad_field = (1u << vbits) - 1u;                    // value field of a padded train row and of an empty entry
    const float base2 = __uint_as_float((uint32_t)(127 + idx_bits - 9) << 23);          // 2^(s-9)
    const float pad_val = __uint_as_float(((uint32_t)(127 + idx_bits - 8) << 23) - 1);   // just under 2^(s-8)
    const float grid = __uint_as_float((uint32_t)(127 - vbits) << 23);                   // 2^-vbits

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(work_head, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total_items) break;
        const int p = item / (mt_per_prob * n_splits);
        const int rem = item - p * (mt_per_prob * n_splits);
        const int mt = rem / n_splits, split = rem - mt * n_splits;
        const TcProblem pr = probs[p];
        const int nq = pool.count[pr.q_slot], nt = tpool.count[pr.t_slot];
        const int q0 = mt * kTM;
        if (q0 >= nq) continue;
        const int n_tiles = (nt + kTN - 1) / kTN;
        const int per = (n_tiles + n_splits - 1) / n_splits;
        const int tile0 = split * per, tile1 = min(n_tiles, tile0 + per);
        // value range of |t|^2 - 2 q.t over the two slots: [-max|q|^2, 2 max|t|^2 + max|q|^2]
        const float nsum = pool.max_norm[pr.q_slot] + tpool.max_norm[pr.t_slot];
        const float off = pool.max_norm[pr.q_slot] + 0.01f * nsum;
        const float sigma = 1.f / (2.04f * nsum + 1e-30f);
        const float m2sigma = -2.f * sigma;

        // the lists hold keys whose index field is RELATIVE to the chunk being scanned and counts downwards
        // (31 - column for the chunk's own columns, more for earlier chunks), so a new key is bits * 2^s + constant
        // with no index register; the three entries are moved along by the chunk distance once per chunk
        uint32_t m1 = pad_field << idx_bits, m2 = m1, m3 = m1;  // empty entries: the value field of a padded row
        uint32_t bnd = 0xffffffffu;                             // QUAD: lower bound of the keys dropped without listing
        int pos = -1;                                            // first split row of the last chunk scanned
        const size_t trow_base = (size_t)pr.t_slot * tpool.slot_rows;
        auto load_train_tile = [&](int tile, int stage) {
            const size_t row0 = trow_base + (size_t)tile * kTN;
            bulk_g2s(smem_u32(sB_hi), tpool.hi + row0 * D, B_BYTES, bar_ld);
            bulk_g2s(smem_u32(sB_lo), tpool.lo + row0 * D, B_BYTES, bar_ld);
            bulk_g2s(smem_u32(sTn + stage * kTN), tpool.norm + row0, kTN * sizeof(float), bar_ld);
        };
        if (tid == 0 && tile0 < tile1) {
            const size_t e0 = ((size_t)pr.q_slot * pool.slot_rows + q0) * D;
            mbar_expect_tx(bar_ld, 2 * A_BYTES + 2 * B_BYTES + kTN * sizeof(float));
            bulk_g2s(smem_u32(sA_hi), pool.hi + e0, A_BYTES, bar_ld);
            bulk_g2s(smem_u32(sA_lo), pool.lo + e0, A_BYTES, bar_ld);
            load_train_tile(tile0, 0);
        }

        for (int tile = tile0; tile < tile1; tile++) {
            const int stage = (tile - tile0) & 1;
            mbar_wait(bar_ld, ph_ld);  // operands of this tile are in shared memory
            ph_ld ^= 1;
            if (tid < kTN) {
                const float tn = sTn[stage * kTN + tid];
                sTnB[stage * kTN + tid] = tn < INFINITY ? fmaf(tn + off, sigma, base2) : pad_val;
            }
            __syncthreads();  // everyone has observed this phase before the next one can be armed; sTnB is written
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = smem_u32(sA_hi), a_lo = smem_u32(sA_lo);
                const uint32_t b_hi = smem_u32(sB_hi), b_lo = smem_u32(sB_lo);
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ks++) {
                    const uint32_t ko = ks * 256;  // two core matrices (16 K-elements) per step
                    umma_bf16(tmem, umma_desc(a_hi + ko, 128, SBO), umma_desc(b_hi + ko, 128, SBO), idesc, ks > 0);
                }
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ks++) {
                    const uint32_t ko = ks * 256;
                    umma_bf16(tmem, umma_desc(a_hi + ko, 128, SBO), umma_desc(b_lo + ko, 128, SBO), idesc, 1);
                    umma_bf16(tmem, umma_desc(a_lo + ko, 128, SBO), umma_desc(b_hi + ko, 128, SBO), idesc, 1);
                }
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma);  // accumulators complete; the train tile in shared memory is free
            ph_mma ^= 1;
            tc_fence_after();
            if (tid == 0 && tile + 1 < tile1) {  // prefetch the next train tile under this tile's epilogue
                mbar_expect_tx(bar_ld, 2 * B_BYTES + kTN * sizeof(float));
                load_train_tile(tile + 1, stage ^ 1);
            }

            // ---- epilogue: warp w reads TMEM lanes 32*(w%4).. (its query rows) and the column quarter w/4
            const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            const float *tn = sTnB + stage * kTN;
            const int cbeg = quarter * (kTN / 4);
            const int lim = min(kTN, nt - tile * kTN);
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + kTN / 4; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + c0, v);
                if (c0 < lim) {  // a chunk of padding rows only: nothing to keep
                    const int idx0 = (tile - tile0) * kTN + c0;
                    const uint32_t delta = pos < 0 ? 0u : (uint32_t)(idx0 - pos);
                    pos = idx0;
                    m1 += delta;
                    m2 += delta;
                    m3 += delta;
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 t4 = *reinterpret_cast<const float4 *>(tn + c0 + c);
                        const uint32_t k0 = __float_as_uint(fmaf(v[c], m2sigma, t4.x)) * key_mul + (31 - c);
                        const uint32_t k1 = __float_as_uint(fmaf(v[c + 1], m2sigma, t4.y)) * key_mul + (30 - c);
                        const uint32_t k2 = __float_as_uint(fmaf(v[c + 2], m2sigma, t4.z)) * key_mul + (29 - c);
                        const uint32_t k3 = __float_as_uint(fmaf(v[c + 3], m2sigma, t4.w)) * key_mul + (28 - c);
                        if (QUAD) {
                            key_insert_quad(m1, m2, m3, bnd, k0, k1, k2, k3);
                        } else {
                            key_insert_pair(m1, m2, m3, k0, k1);
                            key_insert_pair(m1, m2, m3, k2, k3);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();  // every warp has drained TMEM before the next tile's MMA overwrites it
        }
        // ---- merge the four quarter lists of every query row.  The scratch aliases the train-tile buffer: every MMA
        //      that read it has completed (bar_mma), and the next item's bulk copies are issued only after the proxy
        //      fence and the barrier at the loop top.
        uint32_t *s_merge = reinterpret_cast<uint32_t *>(sB_hi);  // [row][quarter][m1, m2, m3, bound]
        const uint32_t idx_mask = (1u << idx_bits) - 1u;
        auto absolute = [&](uint32_t m) {  // relative, descending index field -> row of the split
            return (m & ~idx_mask) | (((uint32_t)(pos + 31) - (m & idx_mask)) & idx_mask);
        };
        *reinterpret_cast<uint4 *>(s_merge + (qrow * 4 + quarter) * 4) = make_uint4(absolute(m1), absolute(m2), absolute(m3), min(m3, bnd));
        __syncthreads();
        if (tid < kTM && q0 + tid < nq) {
            const int row = tid;
            uint32_t best[5] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            uint32_t third = 0xffffffffu;  // smallest of the four threads' bounds (third entry, dropped quad members)
#pragma unroll
            for (int e = 0; e < 16; e++) {
                uint32_t x = s_merge[row * 16 + e];
                if (e % 4 == 3) {
                    third = min(third, x);
                    continue;
                }
#pragma unroll
                for (int r = 0; r < 5; r++) {  // sorted insertion, branch-free
                    const uint32_t lo = min(best[r], x);
                    x = max(best[r], x);
                    best[r] = lo;
                }
            }
            const float inv_sigma = 2.04f * nsum + 1e-30f;
            TcCand o;
#pragma unroll
            for (int r = 0; r < kCand; r++) {
                o.d[r] = INFINITY;
                o.i[r] = -1;
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t field = best[r] >> idx_bits;
                if (field != pad_field) {
                    o.d[r] = (float)field * grid * inv_sigma - off;
                    o.i[r] = tile0 * kTN + (int)(best[r] & idx_mask);
                }
            }
            // lower bound of every value of this split that is not among the four candidates: what the quarter lists
            // dropped (>= their third entries) and what the merge dropped (>= the fifth key), less the quantisation
            // slack (two roundings to the grid of 2^-vbits, the decode's own fp32 roundings)
            const uint32_t bfield = min(third, best[4]) >> idx_bits;
            o.worst = bfield != pad_field ? ((float)bfield - 2.5f) * grid * inv_sigma - off - 2e-6f * nsum : INFINITY;
            cand_out[(size_t)(p * n_splits + split) * cand_stride + q0 + row] = o;
        }
        fence_async_smem();  // generic-proxy accesses to the tile buffer before the next bulk copy overwrites it
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

constexpr int kTP = 128;  // train rows per tile of the warp-specialised kernel

// Warp-specialised form of the key kernel, ONE CTA per SM with the whole TMEM: 16 epilogue warps, an MMA warp and a copy
// warp; no block barrier inside an item.  What the earlier pipelined forms showed (profiles/r2): (1) with two operand
// stages the bulk copy of tile k + 1 can only start when MMA(k - 1) has released its stage and takes longer than MMA(k);
// (2) ONE producer warp doing waits, biased norms, descriptor arithmetic, 12 MMAs and the copies of a tile executes
// ~400 dependent instructions per tile, ~1 us, which is longer than the epilogue of a tile: the epilogue warps waited
// for the tensor core (30 % of the stall samples at the mma_done wait) while the tensor pipe was 21-26 % busy.  Here
//   * NS operand stages (4 for 64-d rows: copies run up to four tiles ahead) and NA = 4 accumulators of 128 TMEM columns;
//   * the MMA warp does nothing but: operands landed, accumulator drained -> 3 * KSTEPS tcgen05.mma (descriptors are
//     one base per operand plus constants) -> commit; the copy warp does the bulk copies and the biased |t|^2;
//   * an epilogue warp gives its part of an accumulator back as soon as tcgen05.ld has put it into registers, BEFORE
//     the key arithmetic, and has the tcgen05.ld of the next tile in flight under the arithmetic of the current one
//     (two register buffers of 32 columns).
// Barriers (mbarrier, parity kept per thread):
//     ld_done[s]   operands of a tile have landed            (bulk copies, expect_tx)        MMA warp, copy warp wait
//     tnb_ready[q] biased |t|^2 of the tile are written      (copy warp, one arrival)        epilogue warps wait
//     mma_done[a]  accumulator complete, operand stage free  (tcgen05.commit)                epilogue warps, copy warp wait
//     drained[a]   accumulator is in registers               (16 arrivals)                   MMA warp waits
// s = tile % NS, a = tile % NA, q = tile % 16.  sTnB has sixteen slots: the copy warp writes slot q of tile k after it
// has seen MMA(k - 1) complete; that MMA was issued after drained(k - 1 - NA), i.e. after every epilogue warp had loaded
// accumulator k - 1 - NA and therefore finished the arithmetic of tile k - 2 - NA >= k - 16.
constexpr int kEpiWarps = 16;
constexpr int kWsThreads = (kEpiWarps + 2) * 32;  // + the MMA warp and the copy warp
constexpr int kWsAcc = 4;       // accumulators (x 128 columns = the SM's 512 TMEM columns)
constexpr int kWsSlots = 16;    // sTnB slots

__host__ __device__ constexpr int ws_stages(int dim) { return dim == 64 ? 4 : 2; }
size_t tc_candidates3w_smem(int dim)
{
    return (size_t)(2 * kTM + 2 * ws_stages(dim) * kTP) * dim * 2 + (ws_stages(dim) + kWsSlots) * kTP * sizeof(float) + 1024;
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// completes every tcgen05.ld of this thread; the registers pass through the asm so that no use is scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

template <int D, bool QUAD>
__global__ void __launch_bounds__(kWsThreads, 1) k_tc_candidates3w(TcPool pool, TcPool tpool, const TcProblem *__restrict__ probs,
                                                         int n_probs, int mt_per_prob, int n_splits, int cand_stride,
                                                         TcCand *__restrict__ cand_out, int *__restrict__ work_head, int idx_bits, uint32_t key_mul)
{
    constexpr int NS = ws_stages(D), NA = kWsAcc, NQ = kWsSlots;
    constexpr int KSTEPS = D / 16;
    constexpr uint32_t SBO = (D / 8) * 128;        // bytes between 8-row groups
    constexpr uint32_t A_BYTES = kTM * D * 2;      // one of hi / lo
    constexpr uint32_t B_BYTES = kTP * D * 2;      // one of hi / lo of one stage
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *sA_hi = sm, *sA_lo = sm + A_BYTES;
    unsigned char *sB = sm + 2 * A_BYTES;          // stage s: hi at sB + s * 2 * B_BYTES, lo right behind it
    float *sTn = reinterpret_cast<float *>(sB + 2 * NS * B_BYTES);  // |t|^2 of the train tiles, one slot per stage (bulk copies)
    float *sTnB = sTn + NS * kTP;                                    // 2^(s-9) + sigma (|t|^2 + off), NQ slots
    __shared__ __align__(8) uint64_t s_bar[2 * NA + NS + NQ];
    __shared__ uint32_t s_tmem;
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: the producer's code stays on the uniform datapath
    const bool producer = warp >= kEpiWarps;
    const int quarter = warp >> 2, qrow = (warp & 3) * 32 + lane;
    const uint32_t bar_mma = smem_u32(&s_bar[0]), bar_dr = smem_u32(&s_bar[NA]);   // + 8 * accumulator
    const uint32_t bar_ld = smem_u32(&s_bar[2 * NA]);                               // + 8 * stage
    const uint32_t bar_tnb = smem_u32(&s_bar[2 * NA + NS]);                         // + 8 * slot
    if (tid == 0) {
        for (int i = 0; i < NA; i++) {
            mbar_init(bar_mma + 8 * i, 1);
            mbar_init(bar_dr + 8 * i, kEpiWarps);
        }
        for (int i = 0; i < NS; i++) mbar_init(bar_ld + 8 * i, 1);
        for (int i = 0; i < NQ; i++) mbar_init(bar_tnb + 8 * i, 1);
        mbar_init_fence();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), NA * kTP);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = umma_idesc(kTM, kTP);
    // parities of the next phase to wait for, one word: bits 0-3 mma_done[a], bits 8-23 ld_done[s] (MMA and copy warp)
    // or tnb_ready[q] (epilogue), bits 16-19 drained[a] (MMA warp)
    uint32_t ph = 0;
    const int total_items = n_probs * mt_per_prob * n_splits;
    const int vbits = 32 - idx_bits;                                  // bits of the quantised value
    const uint32_t pad_field = (1u << vbits) - 1u;                    // value field of a padded train row and of an empty entry
    const uint32_t idx_mask = (1u << idx_bits) - 1u;
    const float base2 = __uint_as_float((uint32_t)(127 + idx_bits - 9) << 23);          // 2^(s-9)
    const float pad_val = __uint_as_float(((uint32_t)(127 + idx_bits - 8) << 23) - 1);   // just under 2^(s-8)
    const float grid = __uint_as_float((uint32_t)(127 - vbits) << 23);                   // 2^-vbits

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(work_head, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total_items) break;
        const int p = item / (mt_per_prob * n_splits);
        const int rem = item - p * (mt_per_prob * n_splits);
        const int mt = rem / n_splits, split = rem - mt * n_splits;
        const TcProblem pr = probs[p];
        const int nq = pool.count[pr.q_slot], nt = tpool.count[pr.t_slot];
        const int q0 = mt * kTM;
        if (q0 >= nq) continue;
        const int n_tiles = (nt + kTP - 1) / kTP;
        const int per = (n_tiles + n_splits - 1) / n_splits;
        const int tile0 = split * per, tile1 = min(n_tiles, tile0 + per);
        const int nk = max(tile1 - tile0, 0);
        // value range of |t|^2 - 2 q.t over the two slots: [-max|q|^2, 2 max|t|^2 + max|q|^2]
        const float nsum = pool.max_norm[pr.q_slot] + tpool.max_norm[pr.t_slot];
        const float off = pool.max_norm[pr.q_slot] + 0.01f * nsum;
        const float sigma = 1.f / (2.04f * nsum + 1e-30f);

        uint32_t m1 = pad_field << idx_bits, m2 = m1, m3 = m1;  // empty entries: the value field of a padded row
        uint32_t bnd = 0xffffffffu;                             // QUAD: lower bound of the keys dropped without listing
        int pos = -1;                                            // first split row of the last chunk scanned
        if (warp == kEpiWarps) {
            // ---- MMA warp: operands landed + accumulator drained -> 3 * KSTEPS tcgen05.mma and the commit
            const uint64_t da_hi = umma_desc(smem_u32(sA_hi), 128, SBO), da_lo = umma_desc(smem_u32(sA_lo), 128, SBO);
            const uint64_t db0 = umma_desc(smem_u32(sB), 128, SBO);
            for (int k = 0; k < nk; k++) {
                const int st = k % NS, ac = k % NA;
                mbar_wait(bar_ld + 8 * st, (ph >> (8 + st)) & 1);  // operands of tile k have landed
                ph ^= 256u << st;
                if (k >= NA) {  // accumulator ac was last used by tile k - NA
                    mbar_wait(bar_dr + 8 * ac, (ph >> (16 + ac)) & 1);
                    ph ^= 65536u << ac;
                }
                // descriptors differ by constants only: +16 (256 bytes) per K step, + st * 2 * B_BYTES per stage
                const uint64_t db_hi = db0 + (uint64_t)((uint32_t)st * (2 * B_BYTES >> 4)), db_lo = db_hi + (B_BYTES >> 4);
                const uint32_t acc = tmem + (uint32_t)ac * kTP;
                if (lane == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        if (ks == 0) umma_bf16_first(acc, da_hi, db_hi, idesc);
                        else umma_bf16_acc(acc, da_hi + 16 * ks, db_hi + 16 * ks, idesc);
                    }
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ks++) {
                        umma_bf16_acc(acc, da_hi + 16 * ks, db_lo + 16 * ks, idesc);
                        umma_bf16_acc(acc, da_lo + 16 * ks, db_hi + 16 * ks, idesc);
                    }
                    umma_commit(bar_mma + 8 * ac);
                }
                __syncwarp();
            }
            for (int k = max(nk - NA, 0); k < nk; k++) {  // every completion is waited for exactly once: the last NA drains
                const int ac = k % NA;
                mbar_wait(bar_dr + 8 * ac, (ph >> (16 + ac)) & 1);
                ph ^= 65536u << ac;
            }
        } else if (warp == kEpiWarps + 1) {
            // ---- copy warp: bulk copies NS tiles ahead, biased |t|^2 of every tile that lands
            const size_t trow_base = (size_t)pr.t_slot * tpool.slot_rows;
            auto load_train_tile = [&](int k) {  // lane 0: arms the stage's barrier and starts the three copies
                const int st = k % NS;
                const size_t row0 = trow_base + (size_t)(tile0 + k) * kTP;
                const uint32_t bar = bar_ld + 8 * st;
                unsigned char *hi = sB + (size_t)st * 2 * B_BYTES;
                mbar_expect_tx(bar, 2 * B_BYTES + kTP * sizeof(float) + (k == 0 ? 2 * A_BYTES : 0));
                bulk_g2s(smem_u32(hi), tpool.hi + row0 * D, B_BYTES, bar);
                bulk_g2s(smem_u32(hi + B_BYTES), tpool.lo + row0 * D, B_BYTES, bar);
                bulk_g2s(smem_u32(sTn + st * kTP), tpool.norm + row0, kTP * sizeof(float), bar);
            };
            if (lane == 0 && nk > 0) {
                const size_t e0 = ((size_t)pr.q_slot * pool.slot_rows + q0) * D;
                load_train_tile(0);  // its barrier also counts the query tile
                bulk_g2s(smem_u32(sA_hi), pool.hi + e0, A_BYTES, bar_ld);
                bulk_g2s(smem_u32(sA_lo), pool.lo + e0, A_BYTES, bar_ld);
                for (int k = 1; k < min(NS, nk); k++) load_train_tile(k);
            }
            for (int k = 0; k < nk; k++) {
                const int st = k % NS, q = k % NQ;
                mbar_wait(bar_ld + 8 * st, (ph >> (8 + st)) & 1);  // |t|^2 of tile k have landed
                ph ^= 256u << st;
                {
                    const float4 t4 = *reinterpret_cast<const float4 *>(sTn + st * kTP + lane * 4);
                    float4 o4;
                    o4.x = t4.x < INFINITY ? fmaf(t4.x + off, sigma, base2) : pad_val;
                    o4.y = t4.y < INFINITY ? fmaf(t4.y + off, sigma, base2) : pad_val;
                    o4.z = t4.z < INFINITY ? fmaf(t4.z + off, sigma, base2) : pad_val;
                    o4.w = t4.w < INFINITY ? fmaf(t4.w + off, sigma, base2) : pad_val;
                    *reinterpret_cast<float4 *>(sTnB + q * kTP + lane * 4) = o4;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tnb + 8 * q);
                const int ao = k % NA;  // MMA(k) complete: its operand stage (and sTn slot, read above) takes tile k + NS
                mbar_wait(bar_mma + 8 * ao, (ph >> ao) & 1);
                ph ^= 1u << ao;
                fence_async_smem();  // the generic reads of sTn[st] above, before the copy that overwrites it
                if (lane == 0 && k + NS < nk) load_train_tile(k + NS);
            }
        } else if (nk > 0) {
            const float m2sigma = -2.f * sigma;
            const int c0 = quarter * (kTP / 4);
            const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
            auto start_load = [&](int k, uint32_t (&r)[32]) {  // accumulator of tile k -> registers (asynchronous)
                const int ac = k % NA;
                mbar_wait(bar_mma + 8 * ac, (ph >> ac) & 1);
                ph ^= 1u << ac;
                tc_fence_after();
                tmem_ld32_nowait(lane_addr + (uint32_t)(ac * kTP), r);
            };
            auto finish_load = [&](int k, uint32_t (&r)[32]) {  // the registers are valid; hand the accumulator back
                tmem_ld_wait(r);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dr + 8 * (k % NA));
            };
            auto keys = [&](int k, const uint32_t (&r)[32]) {
                const int q = k % NQ;
                mbar_wait(bar_tnb + 8 * q, (ph >> (8 + q)) & 1);  // (long since complete: it precedes the MMA's issue)
                ph ^= 256u << q;
                if (c0 < min(kTP, nt - (tile0 + k) * kTP)) {  // else a chunk of padding rows only: nothing to keep
                    const float *tn = sTnB + q * kTP + c0;
                    const int idx0 = k * kTP + c0;
                    const uint32_t delta = pos < 0 ? 0u : (uint32_t)(idx0 - pos);
                    pos = idx0;
                    m1 += delta;
                    m2 += delta;
                    m3 += delta;
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 t4 = *reinterpret_cast<const float4 *>(tn + c);
                        const uint32_t k0 = __float_as_uint(fmaf(__uint_as_float(r[c]), m2sigma, t4.x)) * key_mul + (31 - c);
                        const uint32_t k1 = __float_as_uint(fmaf(__uint_as_float(r[c + 1]), m2sigma, t4.y)) * key_mul + (30 - c);
                        const uint32_t k2 = __float_as_uint(fmaf(__uint_as_float(r[c + 2]), m2sigma, t4.z)) * key_mul + (29 - c);
                        const uint32_t k3 = __float_as_uint(fmaf(__uint_as_float(r[c + 3]), m2sigma, t4.w)) * key_mul + (28 - c);
                        if (QUAD) {
                            key_insert_quad(m1, m2, m3, bnd, k0, k1, k2, k3);
                        } else {
                            key_insert_pair(m1, m2, m3, k0, k1);
                            key_insert_pair(m1, m2, m3, k2, k3);
                        }
                    }
                }
            };
            uint32_t va[32], vb[32];
            start_load(0, va);
            for (int k = 0; k < nk; k += 2) {  // two tiles per trip: the register buffers keep their names
                finish_load(k, va);
                if (k + 1 < nk) start_load(k + 1, vb);
                keys(k, va);
                if (k + 1 < nk) {
                    finish_load(k + 1, vb);
                    if (k + 2 < nk) start_load(k + 2, va);
                    keys(k + 1, vb);
                }
            }
        }
        // ---- merge the four quarter lists of every query row.  The scratch aliases the operand stages: every MMA that
        //      read them has completed, no copy is in flight (the producer has waited for everything it started), and the
        //      next item's bulk copies are issued only after the proxy fence and the barrier at the loop top.
        uint32_t *s_merge = reinterpret_cast<uint32_t *>(sB);  // [row][quarter][m1, m2, m3, bound]
        if (!producer) {  // (an epilogue warp has itself seen the last MMA complete)
            auto absolute = [&](uint32_t m) {  // relative, descending index field -> row of the split
                return (m & ~idx_mask) | (((uint32_t)(pos + 31) - (m & idx_mask)) & idx_mask);
            };
            *reinterpret_cast<uint4 *>(s_merge + (qrow * 4 + quarter) * 4) = make_uint4(absolute(m1), absolute(m2), absolute(m3), min(m3, bnd));
        }
        __syncthreads();
        if (tid < kTM && q0 + tid < nq) {
            const int row = tid;
            uint32_t best[5] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            uint32_t third = 0xffffffffu;  // smallest of the four threads' bounds (third entry, dropped quad members)
#pragma unroll
            for (int e = 0; e < 16; e++) {
                uint32_t x = s_merge[row * 16 + e];
                if (e % 4 == 3) {
                    third = min(third, x);
                    continue;
                }
#pragma unroll
                for (int r = 0; r < 5; r++) {  // sorted insertion, branch-free
                    const uint32_t lo = min(best[r], x);
                    x = max(best[r], x);
                    best[r] = lo;
                }
            }
            const float inv_sigma = 2.04f * nsum + 1e-30f;
            TcCand o;
#pragma unroll
            for (int r = 0; r < kCand; r++) {
                o.d[r] = INFINITY;
                o.i[r] = -1;
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t field = best[r] >> idx_bits;
                if (field != pad_field) {
                    o.d[r] = (float)field * grid * inv_sigma - off;
                    o.i[r] = tile0 * kTP + (int)(best[r] & idx_mask);
                }
            }
            // lower bound of every value of this split that is not among the four candidates (see k_tc_candidates3)
            const uint32_t bfield = min(third, best[4]) >> idx_bits;
            o.worst = bfield != pad_field ? ((float)bfield - 2.5f) * grid * inv_sigma - off - 2e-6f * nsum : INFINITY;
            cand_out[(size_t)(p * n_splits + split) * cand_stride + q0 + row] = o;
        }
        fence_async_smem();  // generic-proxy accesses to the stages before the next bulk copy overwrites them
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, NA * kTP);
}

// ------------------------------------------------------------------------------------------------
// K5c: exact rescoring of the candidates, top-k, margin test.  LPQ lanes per query, one per
// candidate slot of a list (4 when the lists keep 4 entries, else 8).
// ------------------------------------------------------------------------------------------------
template <int D, int LPQ>
__global__ void __launch_bounds__(256) k_tc_rescore(TcPool pool, TcPool tpool, const TcProblem *__restrict__ probs, int n_splits,
                                                    int cand_stride, const TcCand *__restrict__ cand, int k, float ratio,
                                                    surf_dmatch *__restrict__ out, uint8_t *__restrict__ good,
                                                    int *__restrict__ redo_list, int *__restrict__ redo_count)
{
    const int p = blockIdx.y;
    const TcProblem pr = probs[p];
    const int nq = pool.count[pr.q_slot], nt = tpool.count[pr.t_slot];
    const int gq = (blockIdx.x * blockDim.x + threadIdx.x) / LPQ;
    const int lane_c = threadIdx.x % LPQ;
    if (gq >= nq) return;  // LPQ divides the warp: a query's lanes leave together
    const int q_row0 = pr.q_off ? pr.q_off[0] : pr.q_row0;
    const int t_row0 = pr.t_off ? pr.t_off[0] : pr.t_row0;
    const int out_row0 = pr.out_off ? pr.out_off[0] : pr.out_row0;
    const float *qsrc = pr.qsrc, *tsrc = pr.tsrc;
    const float *qv = qsrc + (size_t)(q_row0 + gq) * D;
    const unsigned gmask = ((1u << LPQ) - 1u) << ((threadIdx.x & 31) / LPQ * LPQ);

    // best k over all splits' candidates, by exact distance; each lane owns candidate slot lane_c
    float bd[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    int bi[4] = {-1, -1, -1, -1};
    float worst_kept = INFINITY;  // smallest "LPQ-th approximate value" over the splits
    for (int s = 0; s < n_splits; s++) {
        const TcCand &c = cand[((size_t)p * n_splits + s) * cand_stride + gq];
        const int idx = c.i[lane_c];
        float dist = INFINITY;
        if (idx >= 0) {
            const float4 *tv = reinterpret_cast<const float4 *>(tsrc + (size_t)(t_row0 + idx) * D);
            const float4 *qv4 = reinterpret_cast<const float4 *>(qv);
            float acc = 0.f;
#pragma unroll 4
            for (int d = 0; d < D / 4; d++) {  // same element order as k_match_exact
                const float4 a = __ldg(qv4 + d), b = __ldg(tv + d);
                float df = a.x - b.x; acc = fmaf(df, df, acc);
                df = a.y - b.y; acc = fmaf(df, df, acc);
                df = a.z - b.z; acc = fmaf(df, df, acc);
                df = a.w - b.w; acc = fmaf(df, df, acc);
            }
            dist = acc;
        }
        // a list that is full bounds every train row it did not keep (an unfilled list kept them all)
        if (c.worst < INFINITY) worst_kept = fminf(worst_kept, c.worst);
        // merge this split's LPQ exact distances into the running top-4 (all lanes keep a copy)
#pragma unroll
        for (int r = 0; r < LPQ; r++) {
            const float od = __shfl_sync(gmask, dist, (threadIdx.x & 31) / LPQ * LPQ + r);
            const int oi = __shfl_sync(gmask, idx, (threadIdx.x & 31) / LPQ * LPQ + r);
            if (oi < 0) continue;
            if (od < bd[3] || (od == bd[3] && (unsigned)oi < (unsigned)bi[3])) {
                bd[3] = od;
                bi[3] = oi;
#pragma unroll
                for (int j = 3; j > 0; j--) {
                    const bool sw = bd[j] < bd[j - 1] || (bd[j] == bd[j - 1] && (unsigned)bi[j] < (unsigned)bi[j - 1]);
                    if (sw) {
                        const float td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
                        const int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
                    }
                }
            }
        }
    }
    if (lane_c != 0) return;
    // margin test: every train row outside the candidate lists has approximate value >= worst_kept,
    // hence exact squared distance >= |q|^2 + worst_kept - eps.  If that clears the k-th best exact
    // distance, the top-k is proven; otherwise the exact kernel redoes this query.
    const float qn = pool.norm[(size_t)pr.q_slot * pool.slot_rows + gq];
    const float tn_max = tpool.max_norm[pr.t_slot];
    const float eps = 1.0e-4f * sqrtf(qn * tn_max) + 4.0e-6f * (qn + tn_max) + 1e-30f;
    const int kk = min(k, nt);
    bool proven = true;
    if (worst_kept < INFINITY && kk > 0) proven = (qn + worst_kept - eps) > bd[kk - 1];
    surf_dmatch *o = out + ((size_t)out_row0 + gq) * k;
    for (int r = 0; r < k; r++) {
        o[r].query_idx = gq;
        o[r].train_idx = bi[r];
        o[r].img_idx = bi[r] < 0 ? -1 : 0;
        o[r].distance = sqrtf(bd[r]);
    }
    if (good && k >= 2)
        good[out_row0 + gq] = (bi[0] >= 0 && bi[1] >= 0 && sqrtf(bd[0]) < ratio * sqrtf(bd[1])) ? 1 : 0;
    if (!proven) {  // redo_count[0] = total (reported), redo_count[1 + p] = entries of problem p's list
        atomicAdd(redo_count, 1);
        const int pos = atomicAdd(redo_count + 1 + p, 1);
        int *e = redo_list + 2 * ((size_t)p * cand_stride + pos);
        e[0] = p;
        e[1] = gq;
    }
}

// Exact redo of the queries whose candidate set could not be proven.  A CTA takes NG consecutive entries of
// the redo list; when they belong to one problem (always, for a database query) every train row is loaded once
// and scored against all NG queries, otherwise the entries are scanned one by one.  One train row per thread at
// a time, each distance the same sequential fp32 sum as in k_match_exact, per-thread top-4 per query, then a
// merge across the CTA.  Rare in practice (a handful per 10^5 queries); common only when the train set holds
// many exact duplicates of a descriptor (more tied rows than the candidate lists keep).
template <int D, int NG>
__device__ void redo_scan(const TcProblem &pr, int nt, const int *__restrict__ entries, int ng, int k, float ratio,
                          surf_dmatch *__restrict__ out, uint8_t *__restrict__ good, float (*s_q)[D], float (*s_d)[8][4],
                          int (*s_i)[8][4])
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q_row0 = pr.q_off ? pr.q_off[0] : pr.q_row0;
    const int t_row0 = pr.t_off ? pr.t_off[0] : pr.t_row0;
    const int out_row0 = pr.out_off ? pr.out_off[0] : pr.out_row0;
    __syncthreads();
    for (int i = tid; i < NG * D; i += 256) {
        const int g = i / D, d = i - g * D;
        s_q[g][d] = g < ng ? pr.qsrc[(size_t)(q_row0 + entries[2 * g + 1]) * D + d] : 0.f;
    }
    __syncthreads();
    float bd[NG][4];
    int bi[NG][4];
#pragma unroll
    for (int g = 0; g < NG; g++)
#pragma unroll
        for (int r = 0; r < 4; r++) {
            bd[g][r] = INFINITY;
            bi[g][r] = -1;
        }
    for (int t = tid; t < nt; t += 256) {
        const float4 *tv = reinterpret_cast<const float4 *>(pr.tsrc + (size_t)(t_row0 + t) * D);
        float acc[NG];
#pragma unroll
        for (int g = 0; g < NG; g++) acc[g] = 0.f;
        // the scan is bound by the latency of the row loads: a whole 64-d row (16 loads) is in flight per thread
        float4 row[D / 4 < 16 ? D / 4 : 16];
        constexpr int RB = D / 4 < 16 ? D / 4 : 16;
#pragma unroll
        for (int d0 = 0; d0 < D / 4; d0 += RB) {
#pragma unroll
        for (int d = 0; d < RB; d++) row[d] = __ldg(tv + d0 + d);
#pragma unroll
        for (int dd = 0; dd < RB; dd++) {  // same element order as k_match_exact
            const int d = d0 + dd;
            const float4 b4 = row[dd];
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const float4 a4 = *reinterpret_cast<const float4 *>(&s_q[g][4 * d]);
                float df = a4.x - b4.x; acc[g] = fmaf(df, df, acc[g]);
                df = a4.y - b4.y; acc[g] = fmaf(df, df, acc[g]);
                df = a4.z - b4.z; acc[g] = fmaf(df, df, acc[g]);
                df = a4.w - b4.w; acc[g] = fmaf(df, df, acc[g]);
            }
        }
        }
#pragma unroll
        for (int g = 0; g < NG; g++) {
            if (acc[g] < bd[g][3]) {  // increasing t within a thread: ties keep the lower index
                bd[g][3] = acc[g];
                bi[g][3] = t;
#pragma unroll
                for (int j = 3; j > 0; j--)
                    if (bd[g][j] < bd[g][j - 1]) {
                        const float td = bd[g][j]; bd[g][j] = bd[g][j - 1]; bd[g][j - 1] = td;
                        const int ti = bi[g][j]; bi[g][j] = bi[g][j - 1]; bi[g][j - 1] = ti;
                    }
            }
        }
    }
    // warp merge of 32 sorted lists per query: 4 rounds of arg-min over the list heads (ties -> lower index)
#pragma unroll
    for (int g = 0; g < NG; g++) {
        for (int r = 0; r < 4; r++) {
            float hd = bd[g][0];
            int hi = bi[g][0], src = lane;
#pragma unroll
            for (int s2 = 16; s2 > 0; s2 >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, hd, s2);
                const int oi = __shfl_xor_sync(0xffffffffu, hi, s2);
                const int os = __shfl_xor_sync(0xffffffffu, src, s2);
                const bool take = (oi >= 0) && (hi < 0 || od < hd || (od == hd && oi < hi));
                if (take) { hd = od; hi = oi; src = os; }
            }
            if (lane == src && hi >= 0) {  // pop the winner's head
                bd[g][0] = bd[g][1]; bi[g][0] = bi[g][1]; bd[g][1] = bd[g][2]; bi[g][1] = bi[g][2];
                bd[g][2] = bd[g][3]; bi[g][2] = bi[g][3];
                bd[g][3] = INFINITY; bi[g][3] = -1;
            }
            if (lane == 0) {
                s_d[g][warp][r] = hd;
                s_i[g][warp][r] = hi;
            }
        }
    }
    __syncthreads();
    if (tid < ng) {  // thread g merges the 8 warps' sorted 4-lists of query g
        const int g = tid, gq = entries[2 * g + 1];
        int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        surf_dmatch *o = out + ((size_t)out_row0 + gq) * k;
        float d0 = INFINITY, d1 = INFINITY;
        int i0 = -1, i1 = -1;
        for (int r = 0; r < k; r++) {
            int bw = -1;
            float hd = INFINITY;
            int hi = -1;
            for (int ww = 0; ww < 8; ww++) {
                if (head[ww] >= 4) continue;
                const float od = s_d[g][ww][head[ww]];
                const int oi = s_i[g][ww][head[ww]];
                if (oi < 0) continue;
                if (hi < 0 || od < hd || (od == hd && oi < hi)) { hd = od; hi = oi; bw = ww; }
            }
            if (bw >= 0) head[bw]++;
            o[r].query_idx = gq;
            o[r].train_idx = hi;
            o[r].img_idx = hi < 0 ? -1 : 0;
            o[r].distance = sqrtf(hd);
            if (r == 0) { d0 = hd; i0 = hi; }
            if (r == 1) { d1 = hd; i1 = hi; }
        }
        if (good && k >= 2) good[out_row0 + gq] = (i0 >= 0 && i1 >= 0 && sqrtf(d0) < ratio * sqrtf(d1)) ? 1 : 0;
    }
}

constexpr int kRedoGroup = 8;

// The unproven queries are listed PER PROBLEM (k_tc_rescore), so a CTA takes up to kRedoGroup queries of one problem and
// reads every train row once for all of them: with the three-entry key lists 0.1-0.2 % of a frame's queries are unproven
// (a handful per frame pair), and one scan of the train frame per group is what bounds this kernel.
template <int D>
__global__ void __launch_bounds__(256) k_tc_redo(TcPool tpool, const TcProblem *__restrict__ probs, int list_stride, int k, float ratio,
                                                 surf_dmatch *__restrict__ out, uint8_t *__restrict__ good,
                                                 const int *__restrict__ redo_list, const int *__restrict__ prob_count)
{
    __shared__ float s_d[kRedoGroup][8][4];
    __shared__ int s_i[kRedoGroup][8][4];
    __shared__ __align__(16) float s_q[kRedoGroup][D];
    const int p = blockIdx.y;
    const int n = prob_count[p];
    if (n <= 0) return;
    const TcProblem pr = probs[p];
    const int nt = tpool.count[pr.t_slot];
    const int *list = redo_list + 2 * (size_t)p * list_stride;
    // few entries and CTAs to spare: one query per CTA; else groups of kRedoGroup share every train row.  (Measured: sizing
    // the groups 1 / 2 / 4 to spread a problem's handful of entries over more CTAs is slower, 171 against 134 us -- a scan
    // is bound by the latency of its row loads, not by the flops of the queries that share them.)
    const int group = n <= (int)gridDim.x ? 1 : kRedoGroup;
    for (int w0 = blockIdx.x * group; w0 < n; w0 += gridDim.x * group) {
        const int ng = min(group, n - w0);
        if (group == 1) redo_scan<D, 1>(pr, nt, list + 2 * w0, ng, k, ratio, out, good, s_q, s_d, s_i);
        else redo_scan<D, kRedoGroup>(pr, nt, list + 2 * w0, ng, k, ratio, out, good, s_q, s_d, s_i);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void launch_tc_prep(const TcPool &pool, const TcSegment *segs, int n_segs, int max_rows, int dim, bool reset_norms,
                    cudaStream_t st, int *launches)
{
    if (n_segs <= 0) return;
    if (reset_norms) k_tc_reset<<<(n_segs + 127) / 128, 128, 0, st>>>(pool, segs, n_segs);
    const int G = dim / 8;
    // every row of a segment plus its padding up to the next multiple of 256
    dim3 gprep(((size_t)(max_rows + 2 * kTN) * G + 255) / 256, n_segs);
    if (dim == 64) k_tc_prep<64><<<gprep, 256, 0, st>>>(segs, n_segs, pool);
    else k_tc_prep<128><<<gprep, 256, 0, st>>>(segs, n_segs, pool);
    if (launches) *launches += reset_norms ? 2 : 1;
}

void launch_tc_match(const TcMatchArgs &a, cudaStream_t st, int *launches)
{
    const int D = a.dim;
    const TcPool &tp = a.tpool.hi ? a.tpool : a.pool;
    // ---- prep every segment (rows -> bf16 hi/lo tiles + norms) ----
    launch_tc_prep(a.pool, a.segs, a.n_segs, a.pool.slot_rows, D, true, st, nullptr);
    launch_zero(a.counters, 2 + a.n_probs, st, nullptr);  // [0] work head, [1] redo count, [2 + p] redo entries of problem p

    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = tc_candidates_smem(D);
    if (!attr_set[dev & 63]) {
        cudaFuncSetAttribute(k_tc_candidates<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates_smem(64));
        cudaFuncSetAttribute(k_tc_candidates<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates_smem(64));
        cudaFuncSetAttribute(k_tc_candidates<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates_smem(128));
        cudaFuncSetAttribute(k_tc_candidates<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates_smem(128));
        attr_set[dev & 63] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int mt_per_prob = a.max_q_rows / kTM;
    const bool small_k = a.k <= 2;
    // k <= 2 against train splits of at most 2^14 rows (frame against frame): the branch-free key kernel
    static const bool keys_off = getenv("SURF_TC_NO_KEYS") != nullptr;         // measurement knobs: the filtered-list kernel,
    static const bool keys_serial = getenv("SURF_TC_KEYS_SERIAL") != nullptr;  // the key kernel with MMA and epilogue back to back
    const bool ws = D == 64 && !keys_serial;  // 128-d rows leave room for two operand stages only: serial key kernel
    const int tile_rows = ws ? kTP : kTN;
    const int tiles_cap = (tp.slot_rows + tile_rows - 1) / tile_rows;
    const int rows_per_split = (tiles_cap + a.n_splits - 1) / a.n_splits * tile_rows;
    if (small_k && rows_per_split <= (1 << 14) && !keys_off) {
        int idx_bits = 9;
        while ((1 << idx_bits) < rows_per_split) idx_bits++;
        static bool attr3_set[64] = {};
        if (!attr3_set[dev & 63]) {
            cudaFuncSetAttribute(k_tc_candidates3<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates3_smem(64));
            cudaFuncSetAttribute(k_tc_candidates3<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates3_smem(128));
            cudaFuncSetAttribute(k_tc_candidates3w<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates3w_smem(64));
            cudaFuncSetAttribute(k_tc_candidates3w<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_candidates3w_smem(64));
            attr3_set[dev & 63] = true;
        }
        static const bool keys_pairs = getenv("SURF_TC_KEYS_PAIRS") != nullptr;  // every key listed (no quad tournament)
#define SURF_KEYS_ARGS a.pool, tp, a.probs, a.n_probs, mt_per_prob, a.n_splits, a.max_q_rows, a.cand, a.counters, idx_bits, 1u << idx_bits
        if (ws) {  // one CTA per SM (all 512 TMEM columns)
            if (keys_pairs) k_tc_candidates3w<64, false><<<sms, kWsThreads, tc_candidates3w_smem(64), st>>>(SURF_KEYS_ARGS);
            else k_tc_candidates3w<64, true><<<sms, kWsThreads, tc_candidates3w_smem(64), st>>>(SURF_KEYS_ARGS);
        } else if (D == 64) {
            k_tc_candidates3<64, true><<<2 * sms, kCandThreads, tc_candidates3_smem(64), st>>>(SURF_KEYS_ARGS);
        } else {
            k_tc_candidates3<128, true><<<sms, kCandThreads, tc_candidates3_smem(128), st>>>(SURF_KEYS_ARGS);
        }
#undef SURF_KEYS_ARGS
    } else {
// persistent CTAs: 256 TMEM columns each, so at most two share an SM's 512; the D = 64 / 4-entry variant fits two
// (2 x 101 KB of shared memory, 64 registers x 512 threads), the others run one per SM
#define SURF_TC_LAUNCH(DD, KK)                                                                              \
    k_tc_candidates<DD, KK><<<((DD) == 64 && (KK) <= 4) ? 2 * sms : sms, kCandThreads, smem, st>>>(         \
        a.pool, tp, a.probs, a.n_probs, mt_per_prob, a.n_splits, a.max_q_rows, a.cand, a.counters)
    if (D == 64) {
        if (small_k) SURF_TC_LAUNCH(64, 4); else SURF_TC_LAUNCH(64, 8);
    } else {
        if (small_k) SURF_TC_LAUNCH(128, 4); else SURF_TC_LAUNCH(128, 8);
    }
#undef SURF_TC_LAUNCH
    }
    const int lpq = small_k ? 4 : kCand;  // lists of the small-k candidate kernel hold 4 entries
    dim3 gres(((size_t)a.max_q_rows * lpq + 255) / 256, a.n_probs);
    dim3 gredo(a.n_probs == 1 ? 2 * sms : (2 * sms / a.n_probs > 2 ? 2 * sms / a.n_probs : 2), a.n_probs);
#define SURF_TC_RESCORE(DD, LL)                                                                                        \
    k_tc_rescore<DD, LL><<<gres, 256, 0, st>>>(a.pool, tp, a.probs, a.n_splits, a.max_q_rows, a.cand, a.k, a.ratio, a.out, \
                                               a.good, a.redo_list, a.counters + 1)
    if (D == 64) {
        if (small_k) SURF_TC_RESCORE(64, 4); else SURF_TC_RESCORE(64, kCand);
        k_tc_redo<64><<<gredo, 256, 0, st>>>(tp, a.probs, a.max_q_rows, a.k, a.ratio, a.out, a.good, a.redo_list, a.counters + 2);
    } else {
        if (small_k) SURF_TC_RESCORE(128, 4); else SURF_TC_RESCORE(128, kCand);
        k_tc_redo<128><<<gredo, 256, 0, st>>>(tp, a.probs, a.max_q_rows, a.k, a.ratio, a.out, a.good, a.redo_list, a.counters + 2);
    }
#undef SURF_TC_RESCORE
    *launches += 6;
}

// Multi-image database: global train row -> (image, row within the image).  img_off is ascending with
// img_off[0] = 0; rows of image i are [img_off[i], img_off[i + 1]).
__global__ void k_db_localize(surf_dmatch *__restrict__ m, long n, const int *__restrict__ img_off, int n_img)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = m[i].train_idx;
    if (g < 0) {
        m[i].img_idx = -1;
        return;
    }
    int lo = 0, hi = n_img;  // last image whose first row is <= g
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (img_off[mid] <= g) lo = mid; else hi = mid;
    }
    m[i].img_idx = lo;
    m[i].train_idx = g - img_off[lo];
}

void launch_db_localize(surf_dmatch *m, long n, const int *img_off, int n_img, cudaStream_t st, int *launches)
{
    if (n <= 0) return;
    k_db_localize<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, n, img_off, n_img);
    if (launches) (*launches)++;
}

__global__ void k_tc_stream_tables(TcSegment *segs, TcProblem *probs, int B, int slots, int prev_slot, const float *desc,
                                   const int *off, const float *prev_desc, const int *prev_off)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int slot = (prev_slot + 1 + i) % slots;
    TcSegment sg;
    sg.src = desc;
    sg.off = off + i;
    sg.row0 = 0;
    sg.n = 0;
    sg.slot = slot;
    sg.dst_row0 = 0;
    segs[i] = sg;
    TcProblem p;
    p.q_slot = slot;
    p.t_slot = (prev_slot + i) % slots;
    p.qsrc = desc;
    p.q_off = off + i;
    p.tsrc = i == 0 ? prev_desc : desc;
    p.t_off = i == 0 ? prev_off : off + i - 1;
    p.out_off = off + i;
    p.q_row0 = p.t_row0 = p.out_row0 = 0;
    probs[i] = p;
}

void launch_tc_stream_tables(TcSegment *segs, TcProblem *probs, int B, int slots, int prev_slot, const float *desc,
                             const int *off, const float *prev_desc, const int *prev_off, cudaStream_t st, int *launches)
{
    k_tc_stream_tables<<<(B + 127) / 128, 128, 0, st>>>(segs, probs, B, slots, prev_slot, desc, off, prev_desc, prev_off);
    *launches += 1;
}

__global__ void __launch_bounds__(256) k_copy_rows(const float *__restrict__ src, const int *__restrict__ off, int dim4,
                                                   float *__restrict__ dst, int max_rows, int *__restrict__ dst_off)
{
    const int n = min(off[1] - off[0], max_rows);
    if (dst_off && blockIdx.x == 0 && threadIdx.x == 0) {
        dst_off[0] = 0;
        dst_off[1] = n;
    }
    const float4 *s = reinterpret_cast<const float4 *>(src) + (size_t)off[0] * dim4;
    float4 *d = reinterpret_cast<float4 *>(dst);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n * dim4; i += (size_t)gridDim.x * blockDim.x)
        d[i] = s[i];
}

void launch_copy_rows(const float *src, const int *off, int dim, float *dst, int max_rows, int *dst_off, cudaStream_t st,
                      int *launches)
{
    k_copy_rows<<<148, 256, 0, st>>>(src, off, dim / 4, dst, max_rows, dst_off);
    *launches += 1;
}

}  // namespace surfb200
