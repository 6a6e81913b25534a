// abfit_divergence.cu — observed pairwise methylation divergence + per-sample methylation level.
//
// Replaces DMatrix::from (src/pedigree.rs:213-262) and the valid-site statistics of
// Pedigree::build (src/pedigree.rs:159-183), batched over windows (site segments).
//
// HBM-bound design (17 bytes per sample-site are algorithmic):
//   pass 1  k_pack      streams posteriorMax (f64), rc.meth.lvl (f64) and status (u8) ONCE, fully
//                       coalesced on the site axis, and emits three bit-planes per sample —
//                       valid = post >= thr, t1 = status >= 1, t2 = status >= 2 (thermometer code,
//                       |a-b| = popc(t1a^t1b) + popc(t2a^t2b)) — plus blocked partial sums of
//                       meth_lvl over valid sites.  3 bits/site leave HBM again (2 % of the input).
//   pass 2  k_pairs     all-pairs popcount over the bit-planes staged in shared memory, 4 x 4 sample tiles
//                       in registers (POPC-bound); exact u64
//                       sums, so D = diff / (2 cnt) is bit-identical to the reference for any
//                       sharding of the site axis.
//   pass 3  k_finalize  D = diff/(2 cnt), per-sample methsum / nvalid, p0uu per window.
#include "abfit_internal.h"

namespace abfit {

constexpr int SB_WORDS = 64;           // words (64 sites) per super-block = one warp's share in k_pack
constexpr int64_t EXACT_MAX = 65536;   // windows up to this many sites: sequential (bit-exact) methsum

struct SuperBlock {
    int32_t window;
    int32_t n_words;      // <= SB_WORDS
    int64_t first_word;   // global packed word index
    int64_t first_site;   // site index on the L axis
    int64_t end_site;     // end of the window
};

// pass 1 ------------------------------------------------------------------------------
// 56 registers: two 128-thread blocks fit in the 14 K registers a k_pairs block leaves free on an SM
__global__ void __maxnreg__(56)
k_pack(const uint8_t *__restrict__ status, const double *__restrict__ post, const double *__restrict__ meth,
       int64_t L, int64_t total_words, const SuperBlock *__restrict__ sbs, int n_sb, int sb_first, int sb_end, int S,
       double thr, unsigned long long *__restrict__ V, unsigned long long *__restrict__ T1,
       unsigned long long *__restrict__ T2, double *__restrict__ methpart, long long *__restrict__ nvpart)
{
    // Work unit = one super-block of one sample; units of this launch: super-blocks [sb_first, sb_end) x S samples,
    // unit u = sample u / n_sbc, super-block sb_first + u % n_sbc (consecutive warps read consecutive memory).  A
    // launch either has a warp per unit, or — the overlapped mode of run_divergence, where the pair pass of one chunk
    // of sites runs while the next chunk is being packed — a fixed number of blocks per SM that loop over the units,
    // so that the packing never takes more of an SM than the registers the pair kernel leaves free.
    const int lane = threadIdx.x & 31;
    const int n_sbc = sb_end - sb_first;
    const long long n_units = (long long)n_sbc * S;
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_units; u += n_warps) {
    const int s = (int)(u / n_sbc);
    const int warp = sb_first + (int)(u - (long long)s * n_sbc);
    const SuperBlock sb = sbs[warp];
    const uint8_t *st = status + (size_t)s * L;
    const double *po = post + (size_t)s * L;
    const double *me = meth + (size_t)s * L;
    unsigned long long *Vs = V + (size_t)s * total_words + sb.first_word;
    unsigned long long *T1s = T1 + (size_t)s * total_words + sb.first_word;
    unsigned long long *T2s = T2 + (size_t)s * total_words + sb.first_word;

    double acc = 0.0;  // lane-local, in site order within the lane
    int nv = 0;
    // Lane l owns the two adjacent sites 2l, 2l+1 of every 64-site word, so one 16-byte load per lane and
    // array covers the word (512 contiguous bytes per warp request).  Bit b < 32 of a packed word is site 2b,
    // bit 32 + b is site 2b + 1: any fixed permutation of the sites inside a word is as good as any other for
    // the popcounts of pass 2, as long as every sample and plane uses the same one.
    const bool vec = ((reinterpret_cast<uintptr_t>(po + sb.first_site) | reinterpret_cast<uintptr_t>(me + sb.first_site)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(st + sb.first_site) & 1) == 0;
#pragma unroll 4
    for (int w = 0; w < sb.n_words; ++w) {
        const int64_t i0 = sb.first_site + (int64_t)w * 64 + 2 * lane, i1 = i0 + 1;
        const bool in0 = i0 < sb.end_site, in1 = i1 < sb.end_site;
        double p0 = -1.0, p1 = -1.0, m0 = 0.0, m1 = 0.0;
        int s0 = 0, s1 = 0;
        if (vec && in1) {
            const double2 pp = __ldcs(reinterpret_cast<const double2 *>(po + i0));
            const double2 mm = __ldcs(reinterpret_cast<const double2 *>(me + i0));
            const unsigned short ss = __ldcs(reinterpret_cast<const unsigned short *>(st + i0));
            p0 = pp.x; p1 = pp.y; m0 = mm.x; m1 = mm.y;
            s0 = ss & 0xff; s1 = ss >> 8;
        } else {
            if (in0) { p0 = __ldcs(po + i0); m0 = __ldcs(me + i0); s0 = (int)__ldcs(st + i0); }
            if (in1) { p1 = __ldcs(po + i1); m1 = __ldcs(me + i1); s1 = (int)__ldcs(st + i1); }
        }
        const bool v0 = in0 && (p0 >= thr), v1 = in1 && (p1 >= thr);
        const unsigned bv0 = __ballot_sync(FULL, v0), bv1 = __ballot_sync(FULL, v1);
        const unsigned b10 = __ballot_sync(FULL, s0 >= 1), b11 = __ballot_sync(FULL, s1 >= 1);
        const unsigned b20 = __ballot_sync(FULL, s0 >= 2), b21 = __ballot_sync(FULL, s1 >= 2);
        if (v0) { acc += m0; ++nv; }
        if (v1) { acc += m1; ++nv; }
        if (lane == 0) {
            Vs[w] = (unsigned long long)bv0 | ((unsigned long long)bv1 << 32);
            T1s[w] = (unsigned long long)b10 | ((unsigned long long)b11 << 32);
            T2s[w] = (unsigned long long)b20 | ((unsigned long long)b21 << 32);
        }
    }
    // fixed-shape tree over lanes: deterministic for a given (L, segmentation)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(FULL, acc, o);
        nv += __shfl_down_sync(FULL, nv, o);
    }
    if (lane == 0) {
        methpart[(size_t)s * n_sb + warp] = acc;
        nvpart[(size_t)s * n_sb + warp] = nv;
    }
    }  // units
}

// pass 2 ------------------------------------------------------------------------------
struct PairItem {
    int32_t window;
    int32_t n_words;     // words of this chunk
    int64_t first_word;  // global packed word index
    int32_t single;      // 1: the only chunk of its window -> plain store, else atomicAdd
    int32_t pad;
};

// Register-tiled all-pairs popcount.  A thread owns up to two 4 x 4 tiles of the (i, j) sample matrix (16 pairs
// each, upper triangle incl. the diagonal blocks): per 64-site word it reads the three bit-planes of its 4 row
// samples (6 x 16-byte loads, a broadcast inside a warp, whose lanes share the row block) and of its 4 column
// samples, and updates 16 packed counters (cnt << 16 | diff: an item has at most 256 words, so diff <= 32 768 and
// cnt <= 16 384).  0.75 shared-memory loads per pair and word instead of 6, and one atomic per pair and item of
// up to 256 words instead of one per 34 words: the pass is bound by the POPC pipe, not by the LSU.
// Shared layout: plane[w * Sp + s] (word-major: the samples of a tile are adjacent), Sp = padded S + 2.
constexpr int PAIR_ITEM_WORDS = 256;
constexpr int PAIR_THREADS_MAX = 640;

// 80 registers (instead of the 96 a 640-thread launch bound allows): 51 K of the SM's 64 K registers, so that one
// 256-thread block of k_pack (52 registers) fits beside a block of this kernel and the two passes overlap
__global__ void __maxnreg__(80)
k_pairs(const unsigned long long *__restrict__ V, const unsigned long long *__restrict__ T1,
        const unsigned long long *__restrict__ T2, int64_t total_words, int S, int P,
        const PairItem *__restrict__ items, const ushort2 *__restrict__ tiletab, int n_tiles, int Sp, int sw,
        unsigned long long *__restrict__ diff, unsigned long long *__restrict__ cnt)
{
    extern __shared__ unsigned long long sh[];
    const PairItem it = items[blockIdx.x];
    const size_t plane = (size_t)sw * Sp;
    unsigned long long *sV = sh, *sA = sh + plane, *sB = sh + 2 * plane;
    // rows S .. Sp-1 stay zero: valid = 0, they add nothing
    {
        const int npad = Sp - S;
        for (int q = threadIdx.x; q < sw * npad; q += blockDim.x) {
            const size_t o = (size_t)(q / npad) * Sp + S + (q % npad);
            sV[o] = 0ull;
            sA[o] = 0ull;
            sB[o] = 0ull;
        }
    }
    unsigned long long *dw = diff + (size_t)it.window * P, *cw = cnt + (size_t)it.window * P;
    const int nthr = blockDim.x;
    for (int pass = 0; pass < n_tiles; pass += 2 * nthr) {
        const int ta = pass + threadIdx.x, tb = ta + nthr;
        const bool has_a = ta < n_tiles, has_b = tb < n_tiles;
        const ushort2 A = has_a ? tiletab[ta] : make_ushort2(0, 0), B = has_b ? tiletab[tb] : make_ushort2(0, 0);
        unsigned acc_a[16], acc_b[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc_a[k] = acc_b[k] = 0u;
        for (int w0 = 0; w0 < it.n_words; w0 += sw) {
            const int nw = min(sw, it.n_words - w0);
            __syncthreads();  // the previous sub-chunk has been consumed (and the zero fill is complete)
            // coalesced on the word axis in global memory
            for (int q = threadIdx.x; q < S * nw; q += nthr) {
                const int s = q / nw, w = q - s * nw;
                const size_t g = (size_t)s * total_words + it.first_word + w0 + w;
                sV[(size_t)w * Sp + s] = V[g];
                sA[(size_t)w * Sp + s] = T1[g];
                sB[(size_t)w * Sp + s] = T2[g];
            }
            __syncthreads();
            auto tile = [&](const ushort2 T, unsigned acc[16]) {
                for (int w = 0; w < nw; ++w) {
                    const size_t r = (size_t)w * Sp + 4 * T.x, c = (size_t)w * Sp + 4 * T.y;
                    unsigned long long vi[4], ai[4], bi[4];
                    {
                        const ulonglong2 v0 = *reinterpret_cast<const ulonglong2 *>(sV + r), v1 = *reinterpret_cast<const ulonglong2 *>(sV + r + 2);
                        const ulonglong2 a0 = *reinterpret_cast<const ulonglong2 *>(sA + r), a1 = *reinterpret_cast<const ulonglong2 *>(sA + r + 2);
                        const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(sB + r), b1 = *reinterpret_cast<const ulonglong2 *>(sB + r + 2);
                        vi[0] = v0.x; vi[1] = v0.y; vi[2] = v1.x; vi[3] = v1.y;
                        ai[0] = a0.x; ai[1] = a0.y; ai[2] = a1.x; ai[3] = a1.y;
                        bi[0] = b0.x; bi[1] = b0.y; bi[2] = b1.x; bi[3] = b1.y;
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {  // two column samples per 16-byte load
                        const ulonglong2 vj = *reinterpret_cast<const ulonglong2 *>(sV + c + 2 * h);
                        const ulonglong2 aj = *reinterpret_cast<const ulonglong2 *>(sA + c + 2 * h);
                        const ulonglong2 bj = *reinterpret_cast<const ulonglong2 *>(sB + c + 2 * h);
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const unsigned long long m0 = vi[a] & vj.x, m1 = vi[a] & vj.y;
                            acc[4 * a + 2 * h] += ((unsigned)__popcll(m0) << 16) + (unsigned)__popcll((ai[a] ^ aj.x) & m0) +
                                                  (unsigned)__popcll((bi[a] ^ bj.x) & m0);
                            acc[4 * a + 2 * h + 1] += ((unsigned)__popcll(m1) << 16) + (unsigned)__popcll((ai[a] ^ aj.y) & m1) +
                                                      (unsigned)__popcll((bi[a] ^ bj.y) & m1);
                        }
                    }
                }
            };
            if (has_a) tile(A, acc_a);
            if (has_b) tile(B, acc_b);
        }
        auto flush = [&](const ushort2 T, const unsigned acc[16]) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int i = 4 * T.x + a, j = 4 * T.y + b;
                    if (i < j && j < S) {
                        const size_t p = (size_t)i * S - (size_t)i * (i + 1) / 2 + (size_t)(j - i - 1);
                        const unsigned long long d = acc[4 * a + b] & 0xffffu, c = acc[4 * a + b] >> 16;
                        if (it.single) {
                            dw[p] = d;
                            cw[p] = c;
                        } else {
                            atomicAdd(dw + p, d);
                            atomicAdd(cw + p, c);
                        }
                    }
                }
        };
        if (has_a) flush(A, acc_a);
        if (has_b) flush(B, acc_b);
    }
}

// pass 3 ------------------------------------------------------------------------------
__global__ void k_finalize_pairs(const unsigned long long *__restrict__ diff,
                                 const unsigned long long *__restrict__ cnt, int64_t n, double *__restrict__ D)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) D[i] = (double)diff[i] / (2.0 * (double)cnt[i]);  // src/pedigree.rs:257 (0/0 -> NaN)
}

// per (window, sample): valid-site count and sum of meth_lvl (src/pedigree.rs:159-172)
// One WARP per (window, sample).  Windows up to EXACT_MAX sites are summed in the reference's own order — one
// sequential pass over the window's sites (src/pedigree.rs:171-172) — but read 32 sites at a time, coalesced:
// every lane contributes its site's value (+0.0 when the site is filtered out: x + 0.0 == x, and the sum can never
// be -0.0), the 32 values are exchanged through shared memory and every lane adds them in site order.  (One
// thread per (window, sample) walking its own row thrashed L1: 5.3 ms for 10 000 windows x 1000 sites x 27 samples.)
constexpr int FIN_WARPS = 4;
__global__ void __launch_bounds__(32 * FIN_WARPS)
k_finalize_samples(const double *__restrict__ post, const double *__restrict__ meth, int64_t L,
                   const int64_t *__restrict__ seg, int W, int S, double thr,
                   const double *__restrict__ methpart, const long long *__restrict__ nvpart,
                   const int32_t *__restrict__ sb_first, int n_sb, double *__restrict__ methsum,
                   long long *__restrict__ nvalid)
{
    __shared__ __align__(16) double xch[FIN_WARPS][2][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long idx = (long long)blockIdx.x * FIN_WARPS + wib;
    if (idx >= (long long)W * S) return;  // whole warps leave together
    const int w = (int)(idx / S), s = (int)(idx - (long long)w * S);
    const int64_t a = seg[w], b = seg[w + 1];
    double acc = 0.0;
    long long nv = 0;
    if (b - a <= EXACT_MAX) {
        const double *po = post + (size_t)s * L, *me = meth + (size_t)s * L;
        int buf = 0;
        for (int64_t i0 = a; i0 < b; i0 += 32, buf ^= 1) {
            const int64_t i = i0 + lane;
            const bool valid = i < b && po[i] >= thr;
            const double m = valid ? me[i] : 0.0;
            nv += __popc(__ballot_sync(FULL, valid));
            xch[wib][buf][lane] = m;
            __syncwarp();
            // sites beyond the window's end hold +0.0 as well, so every chunk adds all 32 slots
#pragma unroll
            for (int q = 0; q < 32; q += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(&xch[wib][buf][q]);
                acc += v.x;
                acc += v.y;
            }
            // the other buffer is written next; this one again only after the following __syncwarp
        }
    } else {
        for (int q = sb_first[w]; q < sb_first[w + 1]; ++q) {
            acc += methpart[(size_t)s * n_sb + q];
            nv += nvpart[(size_t)s * n_sb + q];
        }
    }
    if (lane == 0) {
        methsum[idx] = acc;
        nvalid[idx] = nv;
    }
}

// p0uu = mean over samples of (1 - rc_meth_lvl), summed in sample order (src/pedigree.rs:179-183)
__global__ void k_p0uu(const double *__restrict__ methsum, const long long *__restrict__ nvalid, int W, int S,
                       double *__restrict__ p0uu)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        const double rc = methsum[(size_t)w * S + s] / (double)nvalid[(size_t)w * S + s];
        acc += 1.0 - rc;
    }
    p0uu[w] = acc / (double)S;
}

void div_arena_release(DivArena &a)
{
    if (a.p) cudaFree(a.p);
    if (a.st2) cudaStreamDestroy(a.st2);
    if (a.ev_start) cudaEventDestroy(a.ev_start);
    if (a.ev_pack_end) cudaEventDestroy(a.ev_pack_end);
    for (auto &e : a.ev_pack)
        if (e) cudaEventDestroy(e);
    a = DivArena();
}

// -------------------------------------------------------------------------------------
int run_divergence(cudaStream_t st, const uint8_t *d_status, const double *d_post, const double *d_meth, int S,
                   int64_t L, const int64_t *h_seg, int W, double thr, double *d_D, unsigned long long *d_diff,
                   unsigned long long *d_cnt, double *d_methsum, long long *d_nvalid, double *d_p0uu,
                   int *launches, float *ms, DivArena *arena)
{
    // host-side segmentation tables
    std::vector<SuperBlock> sbs;
    std::vector<int32_t> sb_first(W + 1, 0);
    std::vector<int64_t> woff(W + 1, 0);
    for (int w = 0; w < W; ++w) {
        const int64_t len = h_seg[w + 1] - h_seg[w];
        if (len < 0) {
            set_error("seg_offsets must be non-decreasing");
            return ABFIT_ERR_ARG;
        }
        const int64_t nw = (len + 63) / 64;
        woff[w + 1] = woff[w] + nw;
        sb_first[w] = (int32_t)sbs.size();
        for (int64_t f = 0; f < nw; f += SB_WORDS) {
            SuperBlock sb;
            sb.window = w;
            sb.n_words = (int32_t)std::min<int64_t>(SB_WORDS, nw - f);
            sb.first_word = woff[w] + f;
            sb.first_site = h_seg[w] + f * 64;
            sb.end_site = h_seg[w + 1];
            sbs.push_back(sb);
        }
    }
    sb_first[W] = (int32_t)sbs.size();
    const int64_t TW = woff[W];
    const int n_sb = (int)sbs.size();
    const int P = S * (S - 1) / 2;
    *launches = 0;

    // pair pass: items of at most PAIR_ITEM_WORDS words (packed 16-bit counters), staged through shared memory
    // `sw` words at a time; 4 x 4 sample tiles, two per thread
    const int Sp = ((S + 3) & ~3) + 2;  // even (16-byte loads), and w * Sp walks through 8 different bank pairs
    const int nb = (S + 3) / 4;
    const int n_tiles = nb * (nb + 1) / 2;
    const size_t smem_cap = 200 * 1024;
    int sw = (int)std::min<size_t>(32, smem_cap / ((size_t)Sp * 24));
    if (sw < 1) {
        set_error("too many samples for the shared-memory pair kernel");
        return ABFIT_ERR_TOO_LARGE;
    }
    const int pair_threads = std::min(PAIR_THREADS_MAX, std::max(32, (((n_tiles + 1) / 2) + 31) & ~31));
    int n_sm = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
    std::vector<PairItem> items;
    // enough items to give every SM a block, never more than PAIR_ITEM_WORDS words each
    const int64_t chunk_target = std::max<int64_t>(8, std::min<int64_t>(PAIR_ITEM_WORDS, (TW + n_sm - 1) / n_sm));
    for (int w = 0; w < W; ++w) {
        const int64_t nw = woff[w + 1] - woff[w];
        if (nw <= 0) continue;
        int64_t n_chunks = (nw + chunk_target - 1) / chunk_target;
        // one long window (a whole methylome): whole waves of one block per SM
        if (W == 1 && n_chunks > n_sm / 2) n_chunks = ((n_chunks + n_sm - 1) / n_sm) * n_sm;
        const int64_t chunk = (nw + n_chunks - 1) / n_chunks;
        for (int64_t f = 0; f < nw; f += chunk) {
            PairItem it;
            it.window = w;
            it.n_words = (int32_t)std::min<int64_t>(chunk, nw - f);
            it.first_word = woff[w] + f;
            it.single = nw <= chunk;
            it.pad = 0;
            items.push_back(it);
        }
    }
    {
        int32_t longest = 1;
        for (auto &it : items) longest = std::max(longest, it.n_words);
        sw = std::min(sw, (int)longest);  // thousands of one-word windows: a block stages one word, not 32
    }
    std::vector<ushort2> pairtab((size_t)n_tiles);  // tile -> (row block, column block), row-major: a warp shares its rows
    {
        size_t t = 0;
        for (int i = 0; i < nb; ++i)
            for (int j = i; j < nb; ++j) pairtab[t++] = make_ushort2((unsigned short)i, (unsigned short)j);
    }

    // device scratch: one arena, carved at 256-byte boundaries
    unsigned long long *d_V = nullptr, *d_T1 = nullptr, *d_T2 = nullptr;
    double *d_methpart = nullptr;
    long long *d_nvpart = nullptr;
    SuperBlock *d_sbs = nullptr;
    int32_t *d_sbfirst = nullptr;
    int64_t *d_seg = nullptr;
    PairItem *d_items = nullptr;
    ushort2 *d_pairtab = nullptr;
    int rc = 0;
    DivArena local_arena;
    DivArena *ar = arena ? arena : &local_arena;
    auto cleanup = [&]() {
        if (!arena) div_arena_release(local_arena);
    };
#define DV_CUDA(call)                                   \
    do {                                                \
        cudaError_t e_ = (call);                        \
        if (e_ != cudaSuccess) {                        \
            rc = cuda_fail(e_, #call);                  \
            cleanup();                                  \
            return rc;                                  \
        }                                               \
    } while (0)
    const size_t plane = (size_t)S * (size_t)std::max<int64_t>(TW, 1) * 8;
    {
        size_t off = 0;
        auto take = [&](size_t bytes) {
            const size_t o = off;
            off += (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
            return o;
        };
        const size_t o_V = take(plane), o_T1 = take(plane), o_T2 = take(plane);
        const size_t o_mp = take((size_t)S * std::max(n_sb, 1) * 8), o_nv = take((size_t)S * std::max(n_sb, 1) * 8);
        const size_t o_sbs = take(sbs.size() * sizeof(SuperBlock)), o_sbf = take((size_t)(W + 1) * 4);
        const size_t o_seg = take((size_t)(W + 1) * 8), o_it = take(items.size() * sizeof(PairItem));
        const size_t o_pt = take(pairtab.size() * sizeof(ushort2));
        if (off > ar->cap) {
            if (ar->p) {
                DV_CUDA(cudaStreamSynchronize(st));  // earlier work on this stream may still read the old arena
                DV_CUDA(cudaFree(ar->p));
                ar->p = nullptr;
                ar->cap = 0;
            }
            DV_CUDA(cudaMalloc(&ar->p, off));
            ar->cap = off;
        }
        char *base = static_cast<char *>(ar->p);
        d_V = reinterpret_cast<unsigned long long *>(base + o_V);
        d_T1 = reinterpret_cast<unsigned long long *>(base + o_T1);
        d_T2 = reinterpret_cast<unsigned long long *>(base + o_T2);
        d_methpart = reinterpret_cast<double *>(base + o_mp);
        d_nvpart = reinterpret_cast<long long *>(base + o_nv);
        d_sbs = reinterpret_cast<SuperBlock *>(base + o_sbs);
        d_sbfirst = reinterpret_cast<int32_t *>(base + o_sbf);
        d_seg = reinterpret_cast<int64_t *>(base + o_seg);
        d_items = reinterpret_cast<PairItem *>(base + o_it);
        d_pairtab = reinterpret_cast<ushort2 *>(base + o_pt);
    }
    if (!sbs.empty())
        DV_CUDA(cudaMemcpyAsync(d_sbs, sbs.data(), sbs.size() * sizeof(SuperBlock), cudaMemcpyHostToDevice, st));
    DV_CUDA(cudaMemcpyAsync(d_sbfirst, sb_first.data(), (size_t)(W + 1) * 4, cudaMemcpyHostToDevice, st));
    DV_CUDA(cudaMemcpyAsync(d_seg, h_seg, (size_t)(W + 1) * 8, cudaMemcpyHostToDevice, st));
    if (!items.empty())
        DV_CUDA(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(PairItem), cudaMemcpyHostToDevice, st));
    if (!pairtab.empty())
        DV_CUDA(cudaMemcpyAsync(d_pairtab, pairtab.data(), pairtab.size() * sizeof(ushort2),
                                cudaMemcpyHostToDevice, st));
    if (P > 0) {
        DV_CUDA(cudaMemsetAsync(d_diff, 0, (size_t)W * P * 8, st));
        DV_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)W * P * 8, st));
    }

    // ---- pass 1 and pass 2, overlapped ---------------------------------------------------------------------------
    // k_pack is bound by HBM, k_pairs by the POPC pipe: run back to back they leave each other's resource idle (2.9 +
    // 2.5 ms on the C5 shape).  A long site axis is therefore cut into chunks (whole pair items / whole super-blocks):
    // chunk c's pair pass runs on the main stream while chunk c+1 is packed on a second stream, the two kernels sharing
    // every SM (k_pairs: 640 threads x 80 registers, k_pack: 128-thread blocks in the register space that is left).
    int n_chunks = 1;
    // (only whole methylomes of millions of sites: measured on 625 000 sites and on 10 000 windows of 1000 sites the
    // plain sequence is faster)
    if (P > 0 && !items.empty() && n_sb > 0 && W == 1 && TW >= 65536) n_chunks = 4;
    if (const char *e = getenv("ABFIT_DEV_DIV_CHUNKS")) n_chunks = std::max(1, std::min(DivArena::MAX_CHUNKS, atoi(e)));
    n_chunks = (int)std::min<size_t>((size_t)std::max(n_chunks, 1), std::max<size_t>(items.size(), 1));
    if (P == 0 || items.empty() || n_sb == 0) n_chunks = 1;
    cudaStream_t st2 = st;
    cudaEvent_t ev_start = nullptr, ev_pack[DivArena::MAX_CHUNKS] = {};
    if (n_chunks > 1) {
        if (!ar->st2) {
            DV_CUDA(cudaStreamCreateWithFlags(&ar->st2, cudaStreamNonBlocking));
            DV_CUDA(cudaEventCreateWithFlags(&ar->ev_start, cudaEventDisableTiming));
            for (auto &e : ar->ev_pack) DV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            DV_CUDA(cudaEventCreate(&ar->ev_pack_end));
        }
        st2 = ar->st2;
        ev_start = ar->ev_start;
        for (int c = 0; c < n_chunks; ++c) ev_pack[c] = ar->ev_pack[c];
    }
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (ms) {
        for (auto &e : ev) DV_CUDA(cudaEventCreate(&e));
        DV_CUDA(cudaEventRecord(ev[0], st));
    }
    const size_t pair_smem = (size_t)sw * Sp * 24;
    if (P > 0 && !items.empty() && pair_smem > 48 * 1024)
        DV_CUDA(cudaFuncSetAttribute(k_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem));
    if (n_chunks == 1) {
        if (n_sb > 0) {
            const long long units = (long long)n_sb * S;
            k_pack<<<(unsigned)((units + 7) / 8), 256, 0, st>>>(d_status, d_post, d_meth, L, TW, d_sbs, n_sb, 0, n_sb, S, thr,
                                                                d_V, d_T1, d_T2, d_methpart, d_nvpart);
            DV_CUDA(cudaGetLastError());
            ++*launches;
        }
        if (ms) DV_CUDA(cudaEventRecord(ev[1], st));
        if (P > 0 && !items.empty()) {
            k_pairs<<<(unsigned)items.size(), pair_threads, pair_smem, st>>>(d_V, d_T1, d_T2, TW, S, P, d_items, d_pairtab,
                                                                            n_tiles, Sp, sw, d_diff, d_cnt);
            DV_CUDA(cudaGetLastError());
            ++*launches;
        }
    } else {
        // chunk c = items [i0, i1) and every super-block that starts before the chunk's last word
        DV_CUDA(cudaEventRecord(ev_start, st));
        DV_CUDA(cudaStreamWaitEvent(st2, ev_start, 0));
        std::vector<int> item_end(n_chunks), sb_end(n_chunks);
        for (int c = 0; c < n_chunks; ++c) {
            // whole waves of one pair block per SM where the item count allows
            int64_t e = (int64_t)items.size() * (c + 1) / n_chunks;
            if (c + 1 < n_chunks && e > n_sm) e = (e / n_sm) * n_sm;
            item_end[c] = (int)e;
        }
        item_end[n_chunks - 1] = (int)items.size();
        for (int c = 0, q = 0; c < n_chunks; ++c) {
            const int64_t last_word = c + 1 < n_chunks ? items[(size_t)item_end[c]].first_word : TW;
            while (q < n_sb && sbs[(size_t)q].first_word < last_word) ++q;
            sb_end[c] = c + 1 < n_chunks ? q : n_sb;
        }
        // 128-thread blocks (4 warps x 56 registers) slip into the registers a k_pairs block leaves free.  Measured on
        // the C5 shape: a warp per unit 5.02 ms for both passes (4 chunks; 5.40 back to back); a FIXED number of looping
        // blocks per SM, which would guarantee the pair kernel its room, starves the packing instead — it needs
        // ~64 resident warps per SM to stream at HBM speed (2 / 4 blocks per SM: 8.4 / 4.7 ms for the packing alone).
        const int pack_threads = 128;
        long long pack_blocks = 1ll << 40;
        if (const char *e = getenv("ABFIT_DEV_DIV_PACK_BLOCKS")) pack_blocks = (long long)std::max(1, atoi(e)) * n_sm;
        for (int c = 0; c < n_chunks; ++c) {
            const int s0 = c ? sb_end[c - 1] : 0, s1 = sb_end[c];
            if (s1 > s0) {
                const long long units = (long long)(s1 - s0) * S;
                const unsigned grid = (unsigned)std::min<long long>(pack_blocks, (units + pack_threads / 32 - 1) / (pack_threads / 32));
                k_pack<<<grid, pack_threads, 0, st2>>>(d_status, d_post, d_meth, L, TW, d_sbs, n_sb, s0, s1, S, thr, d_V, d_T1, d_T2,
                                                        d_methpart, d_nvpart);
                DV_CUDA(cudaGetLastError());
                ++*launches;
            }
            DV_CUDA(cudaEventRecord(ev_pack[c], st2));
        }
        if (ms) DV_CUDA(cudaEventRecord(ar->ev_pack_end, st2));
        for (int c = 0; c < n_chunks; ++c) {
            const int i0 = c ? item_end[c - 1] : 0, i1 = item_end[c];
            DV_CUDA(cudaStreamWaitEvent(st, ev_pack[c], 0));
            if (i1 > i0) {
                k_pairs<<<(unsigned)(i1 - i0), pair_threads, pair_smem, st>>>(d_V, d_T1, d_T2, TW, S, P, d_items + i0, d_pairtab,
                                                                            n_tiles, Sp, sw, d_diff, d_cnt);
                DV_CUDA(cudaGetLastError());
                ++*launches;
            }
        }
        if (ms) DV_CUDA(cudaEventRecord(ev[1], st));  // end of the pair pass (the packing ended earlier, on st2)
    }
    if (P > 0) {
        const int64_t n = (int64_t)W * P;
        k_finalize_pairs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_diff, d_cnt, n, d_D);
        DV_CUDA(cudaGetLastError());
        ++*launches;
    }
    {
        const int n = W * S;
        k_finalize_samples<<<(n + FIN_WARPS - 1) / FIN_WARPS, 32 * FIN_WARPS, 0, st>>>(
            d_post, d_meth, L, d_seg, W, S, thr, d_methpart, d_nvpart, d_sbfirst, n_sb, d_methsum, d_nvalid);
        DV_CUDA(cudaGetLastError());
        ++*launches;
        if (d_p0uu) {
            k_p0uu<<<(W + 127) / 128, 128, 0, st>>>(d_methsum, d_nvalid, W, S, d_p0uu);
            DV_CUDA(cudaGetLastError());
            ++*launches;
        }
    }
    if (ms) DV_CUDA(cudaEventRecord(ev[2], st));
    DV_CUDA(cudaStreamSynchronize(st));
    if (ms) {
        // one chunk: ms[0] = k_pack (the HBM-bound pass), ms[1] = all-pairs popcount + finalisation.
        // overlapped: ms[0] = until the last chunk was packed, ms[1] = the rest (pair pass of the last chunks +
        // finalisation); ms[0] + ms[1] is the duration of the call's kernels either way
        if (n_chunks > 1) {
            float total = 0.f;
            cudaEventElapsedTime(&ms[0], ev[0], ar->ev_pack_end);
            cudaEventElapsedTime(&total, ev[0], ev[2]);
            ms[1] = total - ms[0];
        } else {
            cudaEventElapsedTime(&ms[0], ev[0], ev[1]);
            cudaEventElapsedTime(&ms[1], ev[1], ev[2]);
        }
        for (auto &e : ev) cudaEventDestroy(e);
    }
    cleanup();
#undef DV_CUDA
    return 0;
}

}  // namespace abfit
