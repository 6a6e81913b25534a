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
//   pass 2  k_pairs     all-pairs popcount over the bit-planes staged in shared memory; exact u64
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
__global__ void __launch_bounds__(256)
k_pack(const uint8_t *__restrict__ status, const double *__restrict__ post, const double *__restrict__ meth,
       int64_t L, int64_t total_words, const SuperBlock *__restrict__ sbs, int n_sb, double thr,
       unsigned long long *__restrict__ V, unsigned long long *__restrict__ T1,
       unsigned long long *__restrict__ T2, double *__restrict__ methpart, long long *__restrict__ nvpart)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int s = blockIdx.y;
    if (warp >= n_sb) return;
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
}

// pass 2 ------------------------------------------------------------------------------
struct PairItem {
    int32_t window;
    int32_t n_words;     // words of this chunk
    int64_t first_word;  // global packed word index
    int32_t single;      // 1: the only chunk of its window -> plain store, else atomicAdd
    int32_t pad;
};

__global__ void __launch_bounds__(256)
k_pairs(const unsigned long long *__restrict__ V, const unsigned long long *__restrict__ T1,
        const unsigned long long *__restrict__ T2, int64_t total_words, int S, int P,
        const PairItem *__restrict__ items, const ushort2 *__restrict__ pairtab,
        unsigned long long *__restrict__ diff, unsigned long long *__restrict__ cnt)
{
    extern __shared__ unsigned long long sh[];
    const PairItem it = items[blockIdx.x];
    const int nw = it.n_words;
    unsigned long long *sV = sh, *sA = sh + (size_t)S * nw, *sB = sh + (size_t)2 * S * nw;
    for (int q = threadIdx.x; q < S * nw; q += blockDim.x) {
        const int s = q / nw, w = q - s * nw;
        const size_t g = (size_t)s * total_words + it.first_word + w;
        sV[q] = V[g];
        sA[q] = T1[g];
        sB[q] = T2[g];
    }
    __syncthreads();
    unsigned long long *dw = diff + (size_t)it.window * P, *cw = cnt + (size_t)it.window * P;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const ushort2 ij = pairtab[p];
        const unsigned long long *vi = sV + (size_t)ij.x * nw, *vj = sV + (size_t)ij.y * nw;
        const unsigned long long *ai = sA + (size_t)ij.x * nw, *aj = sA + (size_t)ij.y * nw;
        const unsigned long long *bi = sB + (size_t)ij.x * nw, *bj = sB + (size_t)ij.y * nw;
        unsigned d = 0, c = 0;  // <= 64 * 2 * nw, nw <= 2^20: fits
        for (int w = 0; w < nw; ++w) {
            const unsigned long long m = vi[w] & vj[w];
            c += __popcll(m);
            d += __popcll((ai[w] ^ aj[w]) & m) + __popcll((bi[w] ^ bj[w]) & m);
        }
        if (it.single) {
            dw[p] = d;
            cw[p] = c;
        } else {
            atomicAdd(dw + p, (unsigned long long)d);
            atomicAdd(cw + p, (unsigned long long)c);
        }
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
__global__ void k_finalize_samples(const double *__restrict__ post, const double *__restrict__ meth, int64_t L,
                                   const int64_t *__restrict__ seg, int W, int S, double thr,
                                   const double *__restrict__ methpart, const long long *__restrict__ nvpart,
                                   const int32_t *__restrict__ sb_first, int n_sb, double *__restrict__ methsum,
                                   long long *__restrict__ nvalid)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= W * S) return;
    const int w = idx / S, s = idx - w * S;
    const int64_t a = seg[w], b = seg[w + 1];
    double acc = 0.0;
    long long nv = 0;
    if (b - a <= EXACT_MAX) {
        // the reference's own order: one sequential pass over the window's sites
        const double *po = post + (size_t)s * L, *me = meth + (size_t)s * L;
        for (int64_t i = a; i < b; ++i)
            if (po[i] >= thr) {
                acc += me[i];
                ++nv;
            }
    } else {
        for (int q = sb_first[w]; q < sb_first[w + 1]; ++q) {
            acc += methpart[(size_t)s * n_sb + q];
            nv += nvpart[(size_t)s * n_sb + q];
        }
    }
    methsum[idx] = acc;
    nvalid[idx] = nv;
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

// -------------------------------------------------------------------------------------
int run_divergence(cudaStream_t st, const uint8_t *d_status, const double *d_post, const double *d_meth, int S,
                   int64_t L, const int64_t *h_seg, int W, double thr, double *d_D, unsigned long long *d_diff,
                   unsigned long long *d_cnt, double *d_methsum, long long *d_nvalid, double *d_p0uu,
                   int *launches, float *ms)
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

    // pair chunks: as many words per block as fit next to S samples in shared memory
    const size_t smem_cap = 160 * 1024;
    int64_t cw_max = (int64_t)(smem_cap / ((size_t)S * 24));
    if (cw_max < 1) {
        set_error("too many samples for the shared-memory pair kernel");
        return ABFIT_ERR_TOO_LARGE;
    }
    if (cw_max > 64) cw_max = 64;
    std::vector<PairItem> items;
    for (int w = 0; w < W; ++w) {
        const int64_t nw = woff[w + 1] - woff[w];
        for (int64_t f = 0; f < nw; f += cw_max) {
            PairItem it;
            it.window = w;
            it.n_words = (int32_t)std::min<int64_t>(cw_max, nw - f);
            it.first_word = woff[w] + f;
            it.single = nw <= cw_max;
            it.pad = 0;
            items.push_back(it);
        }
    }
    std::vector<ushort2> pairtab((size_t)P);
    {
        size_t p = 0;
        for (int i = 0; i < S; ++i)
            for (int j = i + 1; j < S; ++j) pairtab[p++] = make_ushort2((unsigned short)i, (unsigned short)j);
    }

    // device scratch
    unsigned long long *d_V = nullptr, *d_T1 = nullptr, *d_T2 = nullptr;
    double *d_methpart = nullptr;
    long long *d_nvpart = nullptr;
    SuperBlock *d_sbs = nullptr;
    int32_t *d_sbfirst = nullptr;
    int64_t *d_seg = nullptr;
    PairItem *d_items = nullptr;
    ushort2 *d_pairtab = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_V); cudaFree(d_T1); cudaFree(d_T2); cudaFree(d_methpart); cudaFree(d_nvpart);
        cudaFree(d_sbs); cudaFree(d_sbfirst); cudaFree(d_seg); cudaFree(d_items); cudaFree(d_pairtab);
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
    DV_CUDA(cudaMalloc(&d_V, plane));
    DV_CUDA(cudaMalloc(&d_T1, plane));
    DV_CUDA(cudaMalloc(&d_T2, plane));
    DV_CUDA(cudaMalloc(&d_methpart, (size_t)S * std::max(n_sb, 1) * 8));
    DV_CUDA(cudaMalloc(&d_nvpart, (size_t)S * std::max(n_sb, 1) * 8));
    DV_CUDA(cudaMalloc(&d_sbs, std::max<size_t>(sbs.size(), 1) * sizeof(SuperBlock)));
    DV_CUDA(cudaMalloc(&d_sbfirst, (size_t)(W + 1) * 4));
    DV_CUDA(cudaMalloc(&d_seg, (size_t)(W + 1) * 8));
    DV_CUDA(cudaMalloc(&d_items, std::max<size_t>(items.size(), 1) * sizeof(PairItem)));
    DV_CUDA(cudaMalloc(&d_pairtab, std::max<size_t>(pairtab.size(), 1) * sizeof(ushort2)));
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

    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (ms) {
        for (auto &e : ev) DV_CUDA(cudaEventCreate(&e));
        DV_CUDA(cudaEventRecord(ev[0], st));
    }
    if (n_sb > 0) {
        dim3 grid((n_sb + 7) / 8, S);
        k_pack<<<grid, 256, 0, st>>>(d_status, d_post, d_meth, L, TW, d_sbs, n_sb, thr, d_V, d_T1, d_T2,
                                     d_methpart, d_nvpart);
        DV_CUDA(cudaGetLastError());
        ++*launches;
    }
    if (ms) DV_CUDA(cudaEventRecord(ev[1], st));
    if (P > 0 && !items.empty()) {
        const size_t smem = (size_t)S * cw_max * 24;
        if (smem > 48 * 1024)
            DV_CUDA(cudaFuncSetAttribute(k_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pairs<<<(unsigned)items.size(), 256, smem, st>>>(d_V, d_T1, d_T2, TW, S, P, d_items, d_pairtab, d_diff,
                                                          d_cnt);
        DV_CUDA(cudaGetLastError());
        ++*launches;
    }
    if (P > 0) {
        const int64_t n = (int64_t)W * P;
        k_finalize_pairs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_diff, d_cnt, n, d_D);
        DV_CUDA(cudaGetLastError());
        ++*launches;
    }
    {
        const int n = W * S;
        k_finalize_samples<<<(n + 127) / 128, 128, 0, st>>>(d_post, d_meth, L, d_seg, W, S, thr, d_methpart,
                                                           d_nvpart, d_sbfirst, n_sb, d_methsum, d_nvalid);
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
    if (ms) {  // ms[0] = k_pack (the HBM-bound pass), ms[1] = all-pairs popcount + finalisation
        cudaEventElapsedTime(&ms[0], ev[0], ev[1]);
        cudaEventElapsedTime(&ms[1], ev[1], ev[2]);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    cleanup();
#undef DV_CUDA
    return 0;
}

}  // namespace abfit
