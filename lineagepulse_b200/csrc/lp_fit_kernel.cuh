// lp_fit_kernel.cuh -- the fit kernel of the LineagePulse fitting path (sm_100a); the other kernels are in lp_kernels.cuh.
//
//  fit_mudisp_kernel   fitMuDisp / fitMuDispGene / fitMuDispGeneImpulse (fitMeanDispersion.R:336-721,
//                      :853-1039): persistent CTAs pull (gene, start) work items from an atomic queue
//                      and run the whole BFGS for that item; every objective evaluation is a
//                      cooperative pass of the CTA over the gene's count row.
//  eval_ll_kernel      evalLogLikMatrix (evalLogLik.R:307-431)
//  impulse_starts_kernel  initialiseImpulseParameters (initialiseImpulseParameters.R:29-111)
//  pi_* kernels        fitPi (fitDropout.R:418-515) as a lock-step batched BFGS over all cells
//  pack / stats kernels   storage of the count matrix (processSCData.R:260-269 replacement)
//
// Evaluation layout ("lane per point"): the 2n central-difference points of one gradient are
// spread over the lanes of a warp, so a warp handles 32/npad cells at a time and every lane
// evaluates ONE (cell, point) observation-likelihood.  Zero / non-zero divergence then only
// occurs between the few cells that share a warp, every lane keeps a single accumulator, and the
// code is the same for every mean model.  Nothing here is a dense contraction: tensor cores and
// TMEM are deliberately unused; the binding unit is the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include "lp_bfgs.cuh"
#include "lp_model.cuh"
#include "lp_exact.cuh"
#include "lp_eval.cuh"

// Threads of a fit CTA (one CTA works on one (gene, start) item) and CTAs per SM: 32 warps per SM either
// way (register cap 64).  512 x 2 instead of 256 x 4 halves the duration of an item and with it the
// item-length tail of the dynamic queue (balance 0.86 -> 0.96 at 15 000 items, -4 % at 1 184 genes,
// -0.4 % at 5 000); 1024 x 1 balances better still but loses 4 % to the longer barriers.
#ifndef LP_FIT_THREADS
#define LP_FIT_THREADS 512
#endif
#define LP_MAX_WARPS (LP_FIT_THREADS / 32)
#define LP_MAX_CLUSTER 2  // CTAs that evaluate one item together (LP_VW / LP_MAX_WARPS)
#define LP_PI_NP 8        // max drop-out parameters per cell (1 + ofMu + LPB200_MAX_PI_PRED)
#define LP_PI_KP 16       // points per cell per round (2q)
#ifndef LP_FAST_MINBLOCKS
#define LP_FAST_MINBLOCKS 2     // CTAs per SM the specialised fit kernels are compiled for (register cap 64)
#endif
#ifndef LP_GENERIC_MINBLOCKS
#define LP_GENERIC_MINBLOCKS 2
#endif

__device__ __forceinline__ int lp_load_count(const uint16_t *r, int j) {
    unsigned v = r[j];
    return v == LP_NA_U16 ? -1 : (int)v;
}
__device__ __forceinline__ int lp_load_count(const int32_t *r, int j) { return r[j]; }

__device__ __forceinline__ double lp_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Values the compiler must keep (in a register, or spilled) instead of re-deriving them inside the evaluation loop: at the
// 64-register cap it otherwise rebuilds the 64-bit row pointers from (gene, pitch, base) -- nine instructions per gather -- and
// re-reads the thread index (S2R) three times per cell (profiles/README.md R2.12).
#ifndef LP_NO_OPAQUE
__device__ __forceinline__ void lp_keep(int &v) { asm volatile("" : "+r"(v)); }
template <typename T>
__device__ __forceinline__ void lp_keep(const T *&p) { asm volatile("" : "+l"(p)); }
#else
__device__ __forceinline__ void lp_keep(int &) {}
template <typename T>
__device__ __forceinline__ void lp_keep(const T *&) {}
#endif

struct CountView {
    const void *ptr;       // gene-major, row pitch in elements
    size_t pitch;
    const int32_t *xmax;   // per gene
    // per gene: cell indices with the non-zero observations first, then the zeros (NA left out), so that
    // the lanes of a warp take the same zero / non-zero branch except in one boundary warp per pass
    const uint32_t *perm;  // [G][pitch]
    const int32_t *nnz;    // [G] number of non-zero observations
    const int32_t *nobs;   // [G] non-zero + zero observations
};

struct FitArgs {
    DevDesign D;
    DevModel M;
    CountView C;
    const double *drop;      // N x q (device) or nullptr
    const double *pi_glob;   // [N] drop-out rate when it does not depend on the gene (logistic, P == 0)
    const double *l1m_glob;  // [N] log(1 - pi)
    double *pi_scratch;      // [gridDim.x][N] per-CTA rates (logistic, P > 0)
    double *l1m_scratch;
    int n_items, n_starts;
    int *queue;
    const double *theta0;    // [n_items][n_theta]  (fit)   /  nat vectors [G][n_nat] (eval)
    int maxit;
    double reltol, ndeps;
    double *theta_out;       // [n_items][n_theta]
    double *ll_out;
    int *conv_out;
    unsigned *evals_out;
    int use_table;
};

// Shared memory of a fit CTA.  CL: the CTA is one of a thread-block cluster that evaluates an item together;
// the partial sums are then double-buffered (one cluster barrier per evaluation instead of two).
template <bool CL>
struct FitSharedT {
    BfgsState<LPB200_MAX_PARAMS> st;
    PointNat pts[LP_KC];
    double vals[2 * LPB200_MAX_PARAMS];
    double red[CL ? 2 : 1][LP_VW][LP_KC + 1];   // [buffer][virtual warp][point]; +1: conflict-free in both directions
    double tab[LP_NTAB][LP_TAB];
    double tab_phi[LP_NTAB];
    double tab_logphi[LP_NTAB], tab_sphi[LP_NTAB];   // log(phi), stirlerr(phi) of the tables being built
    double tab0_phi;                              // what tab[0] currently holds: dispersion and length (0: nothing)
    int tab0_len;
    int tab_of_point[LP_KC];
    int item;
    unsigned epoch;                               // evaluations done (selects the red buffer)
};
typedef FitSharedT<false> FitShared;

__device__ __forceinline__ unsigned lp_cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned lp_cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void lp_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the double at the same shared-memory offset as `p` in CTA `rank` of the cluster (distributed shared memory)
__device__ __forceinline__ double lp_dsmem_load(const double *p, unsigned rank) {
    unsigned local = (unsigned)__cvta_generic_to_shared(p), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
    return v;
}
__device__ __forceinline__ int lp_dsmem_load_int(const int *p, unsigned rank) {
    unsigned local = (unsigned)__cvta_generic_to_shared(p), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    int v;
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// Evaluation layout.  The cells of a gene (in the order of its permutation: non-zero observations first) are
// dealt to LP_VW = 32 VIRTUAL warps: virtual warp v, slot s takes the cells r = v cpw + s + i 32 cpw
// (cpw = 32 / npad cells per warp iteration, npad = points padded to a power of two), each lane adds its
// (cell, point) terms in that order, the slots of a virtual warp are combined by xor shuffles and the 32
// virtual warps are added in index order.  Which PHYSICAL warp runs a virtual warp does not matter, so the sum
// -- and with it every decision of the optimiser -- is the same for a CTA of 128 or 512 threads and for a
// cluster of 1 or 2 CTAs working on the item together (physical warp p of P takes v = p, p + P, ...).
// The host mirror (tests/csrc/mirror_harness.cpp) restates exactly this order.  (64 virtual warps, which would
// allow clusters of 4, cost 2.4 % at 5 000 x 10 000 -- 12.5 KB more shared memory per CTA -- and clusters of 4
// never paid: profiles/README.md.)
// ---------------------------------------------------------------------------------------------------------
// Lane policies: the (cell, point) term of the specialised and of the generic evaluator.
template <typename CT, int MU, bool ZINB>
struct FastLane {
    FastPoint F;
    const double *tabp, *tt, *sf, *pi_ptr, *l1m_ptr;
    const CT *row;
    int tab_len, nnz;
    template <typename SH>
    __device__ __forceinline__ void setup(const FitArgs &A, SH &S, const CT *row_, int gene, int k, int tab_len_,
                                          const double *pi_ptr_, const double *l1m_ptr_) {
        lp_fast_point<MU>(S.pts[k], F);
        tab_len = tab_len_;
        tabp = tab_len > 0 ? S.tab[S.tab_of_point[k]] : nullptr;
        tt = A.D.t;
        sf = A.D.sf;
        pi_ptr = pi_ptr_;
        l1m_ptr = l1m_ptr_;
        row = row_;
        nnz = A.C.nnz[gene];
    }
    struct Pre {};
    static constexpr bool kDeepPrefetch = false;
    __device__ __forceinline__ double term(int r, int j, const Pre &) const {
        const int x = (r < nnz) ? lp_load_count(row, j) : 0;
        return lp_fast_cell<MU, ZINB>(F, x, j, tt, sf, pi_ptr, l1m_ptr, tabp, tab_len);
    }
    // Gradient request of the impulse model: 16 points = 8 coordinates x (+, -) of one cell in one half warp,
    // point k moves coordinate k / 2 of (phi, beta1, beta2, h0, h1, h2, t1, t2).  Only k = 2, 3, 12, 13 have a first
    // sigmoid of their own and only k = 4, 5, 14, 15 a second one; the other values are those of the unmoved
    // parameters.  Each lane therefore evaluates ONE sigmoid (lane 0 of the half warp the unmoved first one, lane 1
    // the unmoved second one) and fetches the other by shuffle: 16 instead of 32 exp + reciprocal per cell, the
    // same operations on the same operands, hence the same bits as term().  Every lane of the warp must call this.
    static constexpr bool kShareSigmoids = (MU == LP_MU_FAST_IMPULSE);
    // mean of cell j under the lane's point; shared: the 16-point gradient layout above (every lane of the warp calls)
    __device__ __forceinline__ double mean(int j, int lane, bool shared) const {
        if (MU != LP_MU_FAST_IMPULSE) return F.p0;
        const double t = tt[j];
        double s1, s2;
        if (shared) {
            const int k = lane & 15;
            const bool own1 = (k == 2) | (k == 3) | (k == 12) | (k == 13), own2 = (k == 4) | (k == 5) | (k == 14) | (k == 15);
            const bool second = own2 | (k == 1);
            const double a = (second ? F.b2 : -F.p0) * (t - (second ? F.t2 : F.t1));   // term(): -p0 (t - t1), b2 (t - t2)
            const double e = lp_exp(a);
            // e < 1e300 (term()'s test) decided on the upper word: e >= 0 or NaN here, the bound is a little lower, and
            // the checked reciprocal returns the same bits wherever the check-free one applies
#ifndef LP_NO_INTCMP
            const double s = ((unsigned)__double2hiint(e) < 0x7e37e43cu) ? lp_rcp<true>(1.0 + e) : lp_rcp<false>(1.0 + e);
#else
            const double s = (e < 1e300) ? lp_rcp<true>(1.0 + e) : lp_rcp<false>(1.0 + e);
#endif
            const double b1 = __shfl_sync(0xffffffffu, s, lane & 16), b2 = __shfl_sync(0xffffffffu, s, (lane & 16) | 1);
            s1 = own1 ? s : b1;
            s2 = own2 ? s : b2;
        } else {
            lp_sigmoid_pair(-F.p0 * (t - F.t1), F.b2 * (t - F.t2), s1, s2);
        }
        return F.inv_h1 * (F.h0 + (F.h1 - F.h0) * s1) * (F.h2 + (F.h1 - F.h2) * s2);
    }
    __device__ __forceinline__ double tail(double mu, int r, int j) const {
        const int x = (r < nnz) ? lp_load_count(row, j) : 0;
        return lp_fast_cell_mu<ZINB>(F, mu, x, j, sf, pi_ptr, l1m_ptr, tabp, tab_len);
    }
};
// Group means or a spline mean model, mean batch factors, one scalar dispersion per gene, drop-out "none" or
// gene-wise fixed (BASELINE configs C4 / C5).  The point stays in shared memory (its group means / spline
// coefficients / batch factors are indexed by the cell); what the specialisation buys is the pruned code -- the
// operations are lp_cell_ll's, in lp_cell_ll's order, so the bits are the generic evaluator's.
template <typename CT, int MUK, bool ZINB>
struct CovariateLane {
    const DevDesign *D;
    const PointNat *P;
    const double *tabp, *pi_ptr, *l1m_ptr;
    const CT *row;
    double phi, logphi;
    int tab_len, nnz;
    template <typename SH>
    __device__ __forceinline__ void setup(const FitArgs &A, SH &S, const CT *row_, int gene, int k, int tab_len_,
                                          const double *pi_ptr_, const double *l1m_ptr_) {
        D = &A.D;
        P = &S.pts[k];
        phi = P->disp[0];
        logphi = lp_log(phi);
        tab_len = tab_len_;
        tabp = tab_len > 0 ? S.tab[S.tab_of_point[k]] : nullptr;
        pi_ptr = pi_ptr_;
        l1m_ptr = l1m_ptr_;
        row = row_;
        nnz = A.C.nnz[gene];
    }
    // the integer gathers of a cell (count, group index, first batch index) are issued one iteration ahead: they
    // depend on the cell index only, and the chain permutation -> index -> shared-memory parameter is what this
    // evaluator otherwise waits for
    struct Pre {
        int x, g, b0;
    };
    static constexpr bool kDeepPrefetch = true;
    static constexpr bool kShareSigmoids = false;
    __device__ __forceinline__ Pre preload(int r, int j) const {
        Pre p;
        p.x = (r < nnz) ? lp_load_count(row, j) : 0;
        p.g = (MUK == LPB200_MU_GROUPS) ? D->group[j] : 0;   // constant mean with batch factors: MUK == LPB200_MU_CONSTANT
        p.b0 = (D->n_conf_mu > 0) ? D->bidx_mu[j] : 0;
        return p;
    }
    __device__ __forceinline__ double term(int r, int j, const Pre &pre) const {
        const int x = pre.x;
        const int N = D->N;
        double mu;
        if (MUK == LPB200_MU_GROUPS || MUK == LPB200_MU_CONSTANT) {
            mu = P->mu[pre.g];
        } else {
            double lin = 0.0;
            for (int k = 0; k < D->k_mu; k++) lin += D->bs_mu[j + (size_t)k * N] * P->mu[k];
            mu = lp_exp(lin);
        }
        int o = 0;
        for (int c = 0; c < D->n_conf_mu; c++) {
            mu = mu * P->bmu[o + (c == 0 ? pre.b0 : D->bidx_mu[(size_t)c * N + j])];
            o += D->nb_mu[c];
        }
        const double mus = mu * (D->sf ? D->sf[j] : 1.0);
        if (x == 0) return ZINB ? lp_ll_zero_zinb(mus, phi, pi_ptr[j]) : lp_ll_zero_nb(mus, phi, logphi);
        const double xd = (double)x;
        const double xphi = (x < tab_len) ? tabp[x] : lp_nb_xphi_part_cold(xd, phi);
        return lp_ll_nonzero(xd, xphi, mus, phi, ZINB ? l1m_ptr[j] : 0.0);
    }
};

template <typename CT>
struct GenericLane {
    const FitArgs *A;
    const PointNat *P;
    const double *tabp, *pi_ptr, *l1m_ptr;
    const CT *row;
    double phi_s, logphi_s;
    int tab_len, gene;
    template <typename SH>
    __device__ __forceinline__ void setup(const FitArgs &A_, SH &S, const CT *row_, int gene_, int k, int tab_len_,
                                          const double *pi_ptr_, const double *l1m_ptr_) {
        A = &A_;
        P = &S.pts[k];
        phi_s = 1.0;
        logphi_s = 0.0;
        tabp = nullptr;
        tab_len = tab_len_;
        if (lp_disp_is_scalar(A_.D, A_.M)) {
            phi_s = P->disp[0];
            logphi_s = lp_log(phi_s);
            if (tab_len > 0) tabp = S.tab[S.tab_of_point[k]];
        }
        pi_ptr = pi_ptr_;
        l1m_ptr = l1m_ptr_;
        row = row_;
        gene = gene_;
    }
    struct Pre {};
    static constexpr bool kDeepPrefetch = false;
    static constexpr bool kShareSigmoids = false;
    __device__ __forceinline__ double term(int r, int j, const Pre &) const {
        const int x = lp_load_count(row, j);
        const double s = A->D.sf ? A->D.sf[j] : 1.0;
        double pf = 0.0, l1 = 0.0;
        if (pi_ptr) {
            pf = pi_ptr[j];
            l1 = l1m_ptr[j];
        }
        return lp_cell_ll(A->D, A->M, *P, A->drop, gene, j, x, s, pf, l1, phi_s, logphi_s, tabp, tab_len);
    }
};

// Evaluate nk (<= LP_KC) points S.pts[0..nk) for one gene; results in out_vals[0..nk).
// All threads of the CTA (CL: of every CTA of the cluster) must call this.
template <bool CL, typename LANE, typename CT, typename SH>
__device__ void lp_eval_points(const FitArgs &A, SH &S, const CT *row, int gene, int xmax, int nk, int ntabs,
                               bool tables_allowed, double *out_vals, const double *pi_ptr, const double *l1m_ptr) {
    const int tid = threadIdx.x, nwarps = blockDim.x >> 5;
    int lane = tid & 31;
#ifndef LP_NO_OPAQUE
    lane = __shfl_sync(0xffffffffu, lane, lane);   // its own value: a shuffle result is not re-derived from S2R in the loop
#endif
    const unsigned crank = CL ? lp_cluster_ctarank() : 0u, csize = CL ? lp_cluster_nctarank() : 1u;
    const int pwarp = (int)crank * nwarps + (tid >> 5), P = (int)csize * nwarps;   // physical warp of the cluster
    // (x, phi)-only part of dnbinom tabulated per distinct phi when that is cheaper than N x nk
    const int tab_len = tables_allowed ? lp_table_len(A.use_table, ntabs, xmax, nk, A.D.N) : 0;
    if (tab_len > 0) {
        // table 0 (the unmoved dispersion) of a gradient request is the table of the function evaluation the line
        // search accepted just before it: kept when dispersion and length are the same
        const bool keep0 = (S.tab0_len == tab_len) && (S.tab0_phi == S.tab_phi[0]);
        if (tid < ntabs) {
            S.tab_logphi[tid] = lp_log(S.tab_phi[tid]);
            S.tab_sphi[tid] = lp_stirlerr(S.tab_phi[tid]);
        }
        __syncthreads();
        for (int e = tid + (keep0 ? tab_len : 0); e < ntabs * tab_len; e += blockDim.x) {
            int tsel = e / tab_len, x = e - tsel * tab_len;
            S.tab[tsel][x] = (x == 0) ? 0.0 : lp_nb_xphi_part_pre((double)x, S.tab_phi[tsel], S.tab_logphi[tsel], S.tab_sphi[tsel]);
        }
        if (tid == 0) {
            S.tab0_phi = S.tab_phi[0];
            S.tab0_len = tab_len;
        }
    }
    __syncthreads();
    int npad = 1;
    while (npad < nk) npad <<= 1;
    const int cpw = 32 / npad;            // cells per warp per iteration
    const int k = lane & (npad - 1);      // this lane's point
    const int slot = lane / npad;
    const bool active = k < nk;
    const int buf = CL ? (int)(S.epoch & 1u) : 0;
    LANE L;
    lp_keep(row);
    L.setup(A, S, row, gene, active ? k : 0, tab_len, pi_ptr, l1m_ptr);
    const uint32_t *perm = A.C.perm + (size_t)gene * A.C.pitch;
    lp_keep(perm);
    const int nobs = A.C.nobs[gene];
    const int stride = LP_VW * cpw;
    // a lane without a point (k >= nk: point counts that are not powers of two) starts beyond the row and never has a term
    const int r_lane = active ? slot : (1 << 30);
    for (int v = pwarp; v < LP_VW; v += P) {
        double acc = 0.0;
        if constexpr (LANE::kShareSigmoids) {
            // one loop for both requests, warp-uniform trip count (the shuffles of the 16-point pass need every lane)
            const bool shared_pass = (nk == 16);
            int r = v * cpw + r_lane;
            int j = (r < nobs) ? (int)perm[r] : 0;
            for (int r0 = v * cpw; r0 < nobs; r0 += stride, r += stride) {
                const int r_next = r + stride;
                const int j_next = (r_next < nobs) ? (int)perm[r_next] : 0;   // one iteration ahead, as below
                const double mu = L.mean(j, lane, shared_pass);
                if (r < nobs) acc += L.tail(mu, r, j);
                j = j_next;
            }
        } else if (active) {
            int r = v * cpw + slot;
            int j = (r < nobs) ? (int)perm[r] : 0;
            if constexpr (LANE::kDeepPrefetch) {
                // cell index two iterations ahead, the cell's integer gathers one iteration ahead
                int j1 = (r + stride < nobs) ? (int)perm[r + stride] : 0;
                typename LANE::Pre pre = L.preload(r, j);
                for (; r < nobs; r += stride) {
                    const int r2 = r + 2 * stride;
                    const int j2 = (r2 < nobs) ? (int)perm[r2] : 0;
                    const typename LANE::Pre pre1 = L.preload(r + stride, j1);
                    acc += L.term(r, j, pre);
                    j = j1;
                    j1 = j2;
                    pre = pre1;
                }
            } else {
                for (; r < nobs; r += stride) {
                    // the next cell index is fetched one iteration ahead: the gathers of the term (count, pseudotime,
                    // drop-out rate) then hang off a register instead of a second dependent L2 access
                    const int r_next = r + stride;
                    const int j_next = (r_next < nobs) ? (int)perm[r_next] : 0;
                    acc += L.term(r, j, typename LANE::Pre());
                    j = j_next;
                }
            }
        }
        // combine the cell slots of the virtual warp (fixed order => deterministic)
        for (int o = 16; o >= npad; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane < npad && lane < nk) S.red[buf][v][lane] = acc;
    }
    if (CL) {
        // every CTA collects the partial sums of the virtual warps its neighbours ran: one distributed-shared-memory
        // load per thread and value, all in flight together (a serial walk over the remote values costs microseconds)
        lp_cluster_sync();
        for (int e = tid; e < LP_VW * nk; e += blockDim.x) {
            const int v = e / nk, kk = e - v * nk;
            const unsigned owner = (unsigned)((v % P) / nwarps);
            if (owner != crank) S.red[buf][v][kk] = lp_dsmem_load(&S.red[buf][v][kk], owner);
        }
    }
    __syncthreads();
    if (tid < nk) {   // the virtual warps in index order
        double s = 0.0;
#pragma unroll 8
        for (int v = 0; v < LP_VW; v++) s += S.red[buf][v][tid];
        out_vals[tid] = s;
    }
    if (tid == 0) S.epoch++;
    __syncthreads();
}

// drop-out rates that stay fixed while one gene is fitted ("logistic", fitMeanDispersion.R:887-892)
template <typename CT>
__device__ void lp_fixed_pi(const FitArgs &A, int gene, const double *&pi_ptr, const double *&l1m_ptr) {
    const DevDesign &D = A.D;
    pi_ptr = nullptr;
    l1m_ptr = nullptr;
    if (A.M.drop_model != LPB200_DROP_LOGISTIC) return;
    if (D.P == 0) {
        pi_ptr = A.pi_glob;
        l1m_ptr = A.l1m_glob;
        return;
    }
    double *ps = A.pi_scratch + (size_t)blockIdx.x * D.N;
    double *ls = A.l1m_scratch + (size_t)blockIdx.x * D.N;
    for (int j = threadIdx.x; j < D.N; j += blockDim.x) {
        double pi = lp_dropout_rate(lp_drop_lin(D, A.M, A.drop, gene, j, 1.0));
        ps[j] = pi;
        ls[j] = lp_log(1.0 - pi);
    }
    __syncthreads();
    pi_ptr = ps;
    l1m_ptr = ls;
}

// MU = -1: generic evaluator (any model); MU = LP_MU_FAST_*: specialised evaluator (constant or impulse mean, one
// scalar dispersion per gene, no mean confounders, drop-out "none" or gene-wise fixed "logistic": BASELINE
// configs C1-C3; same arithmetic as lp_cell_ll operation for operation -- bit-identical to the generic path --
// but the lane's point lives in registers and the dead model branches are gone; the generic kernel is ~13k SASS
// instructions, far beyond the I-cache), ZINB selects the fixed-drop-out ZINB terms over the NB terms.
// CL: launched as thread-block clusters; the CTAs of a cluster take one (gene, start) item together, each
// running the optimiser state machine redundantly on the same all-cluster sums (item-length tail / cluster size).
template <typename CT, int MU, bool ZINB, bool CL>
__global__ void __launch_bounds__(LP_FIT_THREADS, (MU >= 0) ? LP_FAST_MINBLOCKS : LP_GENERIC_MINBLOCKS)
fit_mudisp_kernel(const __grid_constant__ FitArgs A) {
    extern __shared__ __align__(16) unsigned char lp_smem[];
    typedef FitSharedT<CL> SH;
    SH &S = *reinterpret_cast<SH *>(lp_smem);
    const DevDesign &D = A.D;
    const DevModel &M = A.M;
    const int tid = threadIdx.x;
    const int n = M.n_theta;
    const bool scalar_phi = lp_disp_is_scalar(D, M);
    const unsigned crank = CL ? lp_cluster_ctarank() : 0u;
    if (tid == 0) {
        S.epoch = 0;
        S.tab0_len = 0;
    }

    for (;;) {
        if (tid == 0 && crank == 0) S.item = atomicAdd(A.queue, 1);
        if (CL) {   // the leader's item for every CTA of the cluster (the barrier of the item's first evaluation
            lp_cluster_sync();                                   // orders this read before the leader's next write)
            if (tid == 0 && crank != 0) S.item = lp_dsmem_load_int(&S.item, 0u);
        }
        __syncthreads();
        const int item = S.item;
        if (item >= A.n_items) break;
        const int gene = item / A.n_starts;
        const CT *row = reinterpret_cast<const CT *>(A.C.ptr) + (size_t)gene * A.C.pitch;
        const int xmax = A.C.xmax[gene];
        const double *pi_ptr, *l1m_ptr;
        lp_fixed_pi<CT>(A, gene, pi_ptr, l1m_ptr);
        if (tid == 0) bfgs_start(S.st, n, A.theta0 + (size_t)item * n, A.maxit, A.reltol, A.ndeps);
        __syncthreads();

        for (;;) {
            const int req = S.st.req;
            if (req == LP_REQ_NONE) break;
            const int total = (req == LP_REQ_F) ? 1 : 2 * n;
            for (int base = 0; base < total; base += LP_KC) {
                const int kc = min(LP_KC, total - base);
                if (tid < kc) {
                    const int tsel = lp_request_point(D, M, S.st.b, n, req, base + tid, A.ndeps, scalar_phi, S.pts[tid]);
                    S.tab_of_point[tid] = tsel;
                    if (scalar_phi) S.tab_phi[tsel] = S.pts[tid].disp[0];
                }
                if (tid == LP_KC && scalar_phi && req == LP_REQ_GRAD) {
                    // base phi for the points that do not move the dispersion coordinate
                    S.tab_phi[0] = lp_base_phi(S.st.b);
                }
                __syncthreads();
                const int ntabs = lp_request_ntabs(req, base);
                if constexpr (MU == LP_MU_FAST_CONST || MU == LP_MU_FAST_IMPULSE)
                    lp_eval_points<CL, FastLane<CT, MU, ZINB>>(A, S, row, gene, xmax, kc, ntabs, true, S.vals + base, pi_ptr, l1m_ptr);
                else if constexpr (MU == LP_MU_FAST_GROUPS)
                    lp_eval_points<CL, CovariateLane<CT, LPB200_MU_GROUPS, ZINB>>(A, S, row, gene, xmax, kc, ntabs, true, S.vals + base, pi_ptr, l1m_ptr);
                else if constexpr (MU == LP_MU_FAST_SPLINES)
                    lp_eval_points<CL, CovariateLane<CT, LPB200_MU_SPLINES, ZINB>>(A, S, row, gene, xmax, kc, ntabs, true, S.vals + base, pi_ptr, l1m_ptr);
                else if constexpr (MU == LP_MU_FAST_CONSTCOV)
                    lp_eval_points<CL, CovariateLane<CT, LPB200_MU_CONSTANT, ZINB>>(A, S, row, gene, xmax, kc, ntabs, true, S.vals + base, pi_ptr, l1m_ptr);
                else
                    lp_eval_points<CL, GenericLane<CT>>(A, S, row, gene, xmax, kc, ntabs, scalar_phi, S.vals + base, pi_ptr, l1m_ptr);
            }
            if (tid == 0) bfgs_advance(S.st, S.vals);
            __syncthreads();
        }
        if (crank == 0) {
            if (tid < n) A.theta_out[(size_t)item * n + tid] = S.st.b[tid];
            if (tid == 0) {
                A.ll_out[item] = (S.st.fail == 1001) ? NAN : -S.st.Fmin;
                A.conv_out[item] = S.st.fail;
                A.evals_out[item] = S.st.evals;
            }
        }
        __syncthreads();
    }
    if (CL) lp_cluster_sync();   // no CTA leaves while a neighbour may still read its shared memory
}

