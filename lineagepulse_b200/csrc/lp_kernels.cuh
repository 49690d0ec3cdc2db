// lp_kernels.cuh -- the CUDA kernels of the LineagePulse fitting path other than the fit kernel (lp_fit_kernel.cuh):
// evalLogLikMatrix, impulse starts, the fitPi kernels, fit extraction, count-matrix storage, probes.
#pragma once
#include "lp_fit_kernel.cuh"

// evalLogLikMatrix: one CTA per gene (grid-stride), stored natural-scale parameters
template <typename CT>
__global__ void __launch_bounds__(LP_FIT_THREADS)
eval_ll_kernel(const __grid_constant__ FitArgs A) {
    extern __shared__ __align__(16) unsigned char lp_smem[];
    FitShared &S = *reinterpret_cast<FitShared *>(lp_smem);
    const DevModel &M = A.M;
    const int n_nat = lp_n_nat(M);
    if (threadIdx.x == 0) S.epoch = 0;
    for (int gene = blockIdx.x; gene < A.D.G; gene += gridDim.x) {
        const CT *row = reinterpret_cast<const CT *>(A.C.ptr) + (size_t)gene * A.C.pitch;
        const double *pi_ptr, *l1m_ptr;
        lp_fixed_pi<CT>(A, gene, pi_ptr, l1m_ptr);
        if (threadIdx.x == 0) {
            lp_nat_load(M, A.theta0 + (size_t)gene * n_nat, S.pts[0]);
            S.tab_of_point[0] = 0;
            S.tab_phi[0] = S.pts[0].disp[0];
        }
        __syncthreads();
        lp_eval_points<false, GenericLane<CT>>(A, S, row, gene, A.C.xmax[gene], 1, 1, lp_disp_is_scalar(A.D, M), S.vals, pi_ptr, l1m_ptr);
        if (threadIdx.x == 0) A.ll_out[gene] = S.vals[0];
        __syncthreads();
    }
}

// pi_j and log(1 - pi_j) of the gene-independent logistic model
__global__ void pi_vector_kernel(DevDesign D, DevModel M, const double *drop, double *pi, double *l1m) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D.N) return;
    double p = lp_dropout_rate(lp_drop_lin(D, M, drop, 0, j, 1.0));
    pi[j] = p;
    l1m[j] = lp_log(1.0 - p);
}

// ---------------------------------------------------------------------------------------------
// initialiseImpulseParameters: one CTA per gene.  Q = orthonormal basis of ns(t, 4) (N x 4).
// ---------------------------------------------------------------------------------------------
struct StartsArgs {
    DevDesign D;
    CountView C;
    const double *Q;   // N x 4 column-major
    double tmin, tmax, t_first, t_last;
    double *peak;      // G x 7 column-major
    double *valley;
    double *fit_scratch;  // [gridDim.x][N]
};

struct ArgExt {
    double v;
    int i;
};
__device__ __forceinline__ ArgExt lp_argmax_comb(ArgExt a, ArgExt b) {  // first index wins ties (which.max)
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ ArgExt lp_argmin_comb(ArgExt a, ArgExt b) {
    if (b.v < a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
template <bool IS_MAX>
__device__ ArgExt lp_block_argext(ArgExt a, ArgExt *sh) {
    for (int o = 16; o > 0; o >>= 1) {
        ArgExt b;
        b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
        a = IS_MAX ? lp_argmax_comb(a, b) : lp_argmin_comb(a, b);
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = a;
    __syncthreads();
    ArgExt r = sh[0];
    for (int w = 1; w < nw; w++) r = IS_MAX ? lp_argmax_comb(r, sh[w]) : lp_argmin_comb(r, sh[w]);
    return r;
}
__device__ double lp_block_sum(double v, double *sh) {
    v = lp_warp_sum(v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += sh[w];
    return s;
}

template <typename CT>
__global__ void __launch_bounds__(256) impulse_starts_kernel(const __grid_constant__ StartsArgs A) {
    __shared__ double sh_d[32];
    __shared__ ArgExt sh_a[32];
    __shared__ double sh_out[14];
    const DevDesign &D = A.D;
    const int N = D.N, tid = threadIdx.x;
    double *fit = A.fit_scratch + (size_t)blockIdx.x * N;
    for (int gene = blockIdx.x; gene < D.G; gene += gridDim.x) {
        const CT *row = reinterpret_cast<const CT *>(A.C.ptr) + (size_t)gene * A.C.pitch;
        // c = Q^T y,  y = log(x / s + 1)
        double c[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j = tid; j < N; j += blockDim.x) {
            int x = max(lp_load_count(row, j), 0);   // NA counts as 0 here, as in processSCData.R:262-264
            double s = D.sf ? D.sf[j] : 1.0;
            double y = log((double)x / s + 1.0);
#pragma unroll
            for (int k = 0; k < 4; k++) c[k] += A.Q[j + (size_t)k * N] * y;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) c[k] = lp_block_sum(c[k], sh_d);
        // fitted = exp(Q c); extremes over [1, N), [0, N-1) and [0, N)
        ArgExt mx_t = {-INFINITY, 0x7fffffff}, mn_t = {INFINITY, 0x7fffffff};
        ArgExt mx_h = mx_t, mn_h = mn_t;
        for (int j = tid; j < N; j += blockDim.x) {
            double lin = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) lin += A.Q[j + (size_t)k * N] * c[k];
            double f = exp(lin);
            fit[j] = f;
            ArgExt e = {f, j};
            if (j >= 1) {
                mx_t = lp_argmax_comb(mx_t, e);
                mn_t = lp_argmin_comb(mn_t, e);
            }
            if (j < N - 1) {
                mx_h = lp_argmax_comb(mx_h, e);
                mn_h = lp_argmin_comb(mn_h, e);
            }
        }
        mx_t = lp_block_argext<true>(mx_t, sh_a);
        mn_t = lp_block_argext<false>(mn_t, sh_a);
        mx_h = lp_block_argext<true>(mx_h, sh_a);
        mn_h = lp_block_argext<false>(mn_h, sh_a);
        __syncthreads();
        if (tid == 0) {
            const double *t = D.t;
            double f0 = fit[0], fl = fit[N - 1];
            double fitmax = fmax(mx_t.v, f0), fitmin = fmin(mn_t.v, f0);
            for (int which = 0; which < 2; which++) {
                int i_tail = which == 0 ? mx_t.i : mn_t.i;
                int i_head = which == 0 ? mx_h.i : mn_h.i;
                if (i_tail < 0 || i_tail >= N) i_tail = min(1, N - 1);   // non-finite fits: never index out of range
                if (i_head < 0 || i_head >= N) i_head = 0;
                double T1 = (A.t_first + t[i_tail]) / 2.0;
                double T2 = (t[i_head] + A.t_last) / 2.0;
                if (T1 == A.t_first) {
                    double v = INFINITY;
                    for (int j = 0; j < N; j++)
                        if (t[j] != T1 && t[j] < v) v = t[j];
                    T1 = v;
                }
                if (T2 == A.t_last) {
                    double v = -INFINITY;
                    for (int j = 0; j < N; j++)
                        if (t[j] != T2 && t[j] > v) v = t[j];
                    if (which == 0) T1 = v;  // initialiseImpulseParameters.R:65-69 assigns scaT1Peak
                    else T2 = v;
                }
                double *o = sh_out + which * 7;
                o[0] = 2.0 * log(3.0) / (T1 - A.tmin);
                o[1] = 2.0 * log(3.0) / (A.tmax - T2);
                o[2] = log(f0 + 1.0);
                o[3] = log((which == 0 ? fitmax : fitmin) + 1.0);
                o[4] = log(fl + 1.0);
                o[5] = T1;
                o[6] = T2;
            }
        }
        __syncthreads();
        if (tid < 7) {
            A.peak[gene + (size_t)tid * D.G] = sh_out[tid];
            A.valley[gene + (size_t)tid * D.G] = sh_out[7 + tid];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// fitPi: lock-step batched BFGS over cells.  Round = pi_eval (partial sums over this shard's
// genes) -> pi_reduce -> [all-reduce across gene shards] -> pi_advance.
// ---------------------------------------------------------------------------------------------
typedef BfgsState<LP_PI_NP> PiState;

struct PiArgs {
    DevDesign D;
    DevModel M;
    CountView C;
    const double *nat;      // [G][n_nat] natural-scale gene parameters
    PiState *states;        // [N]
    int kp;                 // live point columns of a round: 2q (buffers below hold kp columns, not LP_PI_KP)
    long long *partial_i;   // [3][n_chunks][kp][N]: exact (a, b, bad) sums of the chunk's 16-gene blocks
    long long *vals_i;      // [3][kp][N]: exact (a, b, bad) sums over blocks -- the all-reduced block
    double *vals;           // [kp][N]
    // tail of a sharded run: the remaining cells' counts over ALL genes (gathered from every shard), cell-major
    const int32_t *tail_counts;   // [n_tail][tail_ld] or nullptr (then the tail reads the resident matrix)
    size_t tail_ld;
    // (x, phi)-only part of dnbinom per gene, tabulated once per fitPi call when the dispersion is one scalar per gene
    // (the gene parameters do not change during the call; every round of every cell would recompute it otherwise)
    const double *xphi_tab;       // [G][xphi_w] or nullptr
    int xphi_w;
    int genes_per_chunk, n_chunks;
    int shared_nk;          // > 0: ForAllCells, every cell evaluates shared_theta[0..shared_nk)
    double shared_theta[LP_PI_KP * LP_PI_NP];  // optimiser coordinates
    int *n_active;
    double ndeps;
};

// theta (optimiser coordinates) -> linear model (fitDropout.R:72-81)
LP_HD void lp_pi_transform(int drop_model, int q, const double *th, double *lin) {
    for (int k = 0; k < q; k++) lin[k] = th[k];
    if (drop_model == LPB200_DROP_LOGISTIC_OFMU) lin[1] = -exp(lin[1]);
    for (int k = 0; k < q; k++) lin[k] = lp_clamp(lin[k], -LP_CLAMP_HI, LP_CLAMP_HI);
    if (drop_model == LPB200_DROP_LOGISTIC_OFMU)
        if (lin[1] > -LP_CLAMP_LO) lin[1] = -LP_CLAMP_LO;
}

#define LP_PI_TILE LP_PI_BLOCK   // genes staged in shared memory per step = one summation block

// tab[g][x] = lp_nb_xphi_part(x, phi_g) for x = 1 .. min(xmax_g, w - 1)
__global__ void __launch_bounds__(256) pi_xphi_table_kernel(const double *nat, int n_nat, const int32_t *xmax, int w, double *tab) {
    const int g = blockIdx.x;
    const double phi = nat[(size_t)g * n_nat];   // nat vector = [disp | ...]
    const int top = min(xmax[g], w - 1);
    for (int x = threadIdx.x; x <= top; x += blockDim.x) tab[(size_t)g * w + x] = (x == 0) ? 0.0 : lp_nb_xphi_part((double)x, phi);
}

// HOIST: gene-independent logistic drop-out model (no matPiConstPredictors, not logistic_ofMu): the rate and
// log(1 - rate) of each of the cell's points are computed once per cell instead of once per (gene, cell, point) --
// the same operations on the same operands, hence the same bits.
// KPMAX: compile-time bound of the points per cell (2q): 2 for the plain logistic model -- the per-point arrays of a
// 16-point build cost 248 registers per thread (8 warps per SM); 2 / 4 / 16 are instantiated.
template <typename CT, bool HOIST, int KPMAX>
__global__ void __launch_bounds__(128) pi_eval_kernel(const __grid_constant__ PiArgs A) {
    __shared__ PointNat sg[LP_PI_TILE];
    const DevDesign &D = A.D;
    const DevModel &M = A.M;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = M.q_drop;
    const int n_nat = lp_n_nat(M);
    const int g0 = blockIdx.y * A.genes_per_chunk;
    const int g1 = min(D.G, g0 + A.genes_per_chunk);

    // this cell's evaluation points
    int nk = 0;
    double coef[KPMAX][LP_PI_NP];
    if (j < D.N) {
        if (A.shared_nk > 0) {
            nk = min(A.shared_nk, KPMAX);   // (the launcher picks KPMAX >= 2q)
            for (int k = 0; k < nk; k++) lp_pi_transform(M.drop_model, q, A.shared_theta + k * LP_PI_NP, coef[k]);
        } else {
            const PiState &st = A.states[j];
            int req = st.req;
            nk = min((req == LP_REQ_F) ? 1 : (req == LP_REQ_GRAD ? 2 * q : 0), KPMAX);
            for (int k = 0; k < nk; k++) {
                double th[LP_PI_NP];
                for (int i = 0; i < q; i++) th[i] = st.b[i];
                if (req == LP_REQ_GRAD) {
                    int c = bfgs_grad_coord(k);
                    th[c] = bfgs_grad_point(st.b[c], k, A.ndeps);
                }
                lp_pi_transform(M.drop_model, q, th, coef[k]);
            }
        }
    }
    // a block of LP_PI_BLOCK genes is summed in gene order in double; blocks are combined exactly
    ExactSum ex[KPMAX];
    double acc[KPMAX];
#pragma unroll
    for (int k = 0; k < KPMAX; k++) {
        acc[k] = 0.0;
        ex[k].a = ex[k].b = ex[k].bad = 0;
    }
    const double s = (j < D.N && D.sf) ? D.sf[j] : 1.0;
    const CT *base = reinterpret_cast<const CT *>(A.C.ptr);
    double pi_h[KPMAX], l1m_h[KPMAX];
    if (HOIST) {
#pragma unroll
        for (int k = 0; k < KPMAX; k++) {
            if (k < nk) {
                pi_h[k] = lp_dropout_rate(1.0 * coef[k][0]);
                l1m_h[k] = log(1.0 - pi_h[k]);
            }
        }
    }
    const double *xtab = A.xphi_tab;
    const int xw = A.xphi_w;

    for (int gt = g0; gt < g1; gt += LP_PI_TILE) {
        __syncthreads();
        if (threadIdx.x < LP_PI_TILE && gt + threadIdx.x < g1)
            lp_nat_load(M, A.nat + (size_t)(gt + threadIdx.x) * n_nat, sg[threadIdx.x]);
        __syncthreads();
        if (nk == 0) continue;
        const int ge = min(LP_PI_TILE, g1 - gt);
        for (int gi = 0; gi < ge; gi++) {
            const int i = gt + gi;
            const int x = lp_load_count(base + (size_t)i * A.C.pitch, j);
            if (x < 0) continue;
            const PointNat &P = sg[gi];
            const double mu = lp_mu_at(D, M, P, j);
            const double phi = lp_disp_at(D, M, P, j);
            const double mus = mu * s;
            const double lmu = (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) ? log(mu) : 0.0;
            double core;  // theta-independent part
            if (x == 0) core = exp(phi * log(phi / (phi + mus)));
            else if (lp_nb_small_x((double)x, phi)) core = lp_dnbinom_small_x((double)x, mus, phi);
            else core = ((xtab && x < xw) ? xtab[(size_t)i * xw + x] : lp_nb_xphi_part((double)x, phi)) + lp_nb_mu_part((double)x, mus, phi);
#pragma unroll
            for (int k = 0; k < KPMAX; k++) {
                if (k < nk) {
                    double v;
                    if (HOIST) {
                        v = (x == 0) ? log((1.0 - pi_h[k]) * core + pi_h[k]) : l1m_h[k] + core;
                    } else {
                        int u = 1;
                        double lin = 1.0 * coef[k][0];
                        if (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) lin += lmu * coef[k][u++];
                        for (int p = 0; p < D.P; p++, u++) lin += D.pi_pred[i + (size_t)p * D.G] * coef[k][u];
                        double pi = lp_dropout_rate(lin);
                        v = (x == 0) ? log((1.0 - pi) * core + pi) : log(1.0 - pi) + core;
                    }
                    if (v < LP_LL_FLOOR) v = LP_LL_FLOOR;
                    acc[k] += v;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KPMAX; k++) {
            if (k < nk) {
                lp_exact_add(ex[k], acc[k]);
                acc[k] = 0.0;
            }
        }
    }
    if (j < D.N) {
        const size_t plane = (size_t)A.n_chunks * A.kp * D.N;
#pragma unroll
        for (int k = 0; k < KPMAX; k++) {
            if (k < nk) {
                const size_t o = ((size_t)blockIdx.y * A.kp + k) * D.N + j;
                A.partial_i[o] = ex[k].a;
                A.partial_i[plane + o] = ex[k].b;
                A.partial_i[2 * plane + o] = ex[k].bad;
            }
        }
    }
}

// vals_i[.][k][j] = integer sum over this shard's gene chunks
__global__ void pi_reduce_kernel(PiArgs A) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.D.N) return;
    int nk;
    if (A.shared_nk > 0) nk = A.shared_nk;
    else {
        int req = A.states[j].req;
        nk = (req == LP_REQ_F) ? 1 : (req == LP_REQ_GRAD ? 2 * A.M.q_drop : 0);
    }
    const size_t plane = (size_t)A.kp * A.D.N;
    for (int k = 0; k < A.kp; k++) {
        ExactSum s = {0, 0, 0};
        if (k < nk) {
            const size_t cplane = (size_t)A.n_chunks * A.kp * A.D.N;
            for (int c = 0; c < A.n_chunks; c++) {
                const size_t oc = ((size_t)c * A.kp + k) * A.D.N + j;
                s.a += A.partial_i[oc];
                s.b += A.partial_i[cplane + oc];
                s.bad += A.partial_i[2 * cplane + oc];
            }
        }
        const size_t o = (size_t)k * A.D.N + j;   // inactive slots are written as 0 so the all-reduce is well defined
        A.vals_i[o] = s.a;
        A.vals_i[plane + o] = s.b;
        A.vals_i[2 * plane + o] = s.bad;
    }
}

// after the all-reduce: back to doubles
__global__ void pi_combine_kernel(PiArgs A) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.D.N) return;
    const size_t plane = (size_t)A.kp * A.D.N;
    for (int k = 0; k < A.kp; k++) {
        const size_t o = (size_t)k * A.D.N + j;
        A.vals[o] = lp_exact_value(A.vals_i[o], A.vals_i[plane + o], A.vals_i[2 * plane + o]);
    }
}

__global__ void pi_start_kernel(PiArgs A, const double *drop, int maxit, double reltol) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.D.N) return;
    const int q = A.M.q_drop;
    double th[LP_PI_NP];
    for (int k = 0; k < q; k++) th[k] = drop[j + (size_t)k * A.D.N];
    if (A.M.drop_model == LPB200_DROP_LOGISTIC_OFMU) th[1] = log(-th[1]);  // fitDropout.R:463-466
    bfgs_start(A.states[j], q, th, maxit, reltol, A.ndeps);
}

__global__ void pi_advance_kernel(PiArgs A) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.D.N) return;
    PiState &st = A.states[j];
    if (st.req == LP_REQ_NONE) return;
    double v[LP_PI_KP];
    for (int k = 0; k < A.kp; k++) v[k] = A.vals[(size_t)k * A.D.N + j];
    bfgs_advance(st, v);
    if (st.req != LP_REQ_NONE) atomicAdd(A.n_active, 1);
}

__global__ void pi_finish_kernel(PiArgs A, double *drop_out, double *ll, int *conv, unsigned *evals) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.D.N) return;
    const PiState &st = A.states[j];
    const int q = A.M.q_drop;
    double lin[LP_PI_NP];
    lp_pi_transform(A.M.drop_model, q, st.b, lin);  // fitDropout.R:302-311
    if (st.fail != 1001)
        for (int k = 0; k < q; k++) drop_out[j + (size_t)k * A.D.N] = lin[k];
    ll[j] = (st.fail == 1001) ? NAN : -st.Fmin;
    conv[j] = st.fail;
    evals[j] = st.evals;
}

// Tail of the lock-step loop (single GPU): the few cells that still iterate are finished by one
// CTA each, which runs the remaining BFGS iterations inside the kernel (no launch / host round
// trip per objective evaluation).  The CTA strides over the genes of its cell; reads of the
// gene-major matrix are strided, which is irrelevant for a handful of cells.
// the still-active cells in cell order (one CTA: the list must be the same on every shard of a sharded run)
__global__ void __launch_bounds__(256) pi_compact_kernel(PiArgs A, int *list, int *count) {
    __shared__ int warp_cnt[8];
    __shared__ int run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) run = 0;
    __syncthreads();
    for (int base = 0; base < A.D.N; base += 256) {
        const int j = base + tid;
        const bool act = j < A.D.N && A.states[j].req != LP_REQ_NONE;
        const unsigned b = __ballot_sync(0xffffffffu, act);
        if (lane == 0) warp_cnt[warp] = __popc(b);
        __syncthreads();
        int off = run;
        for (int w = 0; w < warp; w++) off += warp_cnt[w];
        if (act) list[off + __popc(b & ((1u << lane) - 1u))] = j;
        __syncthreads();
        if (tid == 0)
            for (int w = 0; w < 8; w++) run += warp_cnt[w];
        __syncthreads();
    }
    if (tid == 0) *count = run;
}

template <typename CT>
__global__ void __launch_bounds__(256) pi_tail_kernel(const __grid_constant__ PiArgs A, const int *list) {
    __shared__ double s_vals[LP_PI_KP];
    __shared__ long long s_ex[LP_PI_KP][3];
    __shared__ double s_coef[LP_PI_KP][LP_PI_NP];
    __shared__ double s_pi[LP_PI_KP], s_l1m[LP_PI_KP];   // gene-independent logistic model: rate and log(1 - rate) per point
    __shared__ int s_nk;
    const DevDesign &D = A.D;
    const DevModel &M = A.M;
    const int j = list[blockIdx.x];
    const int q = M.q_drop, n_nat = lp_n_nat(M);
    const int tid = threadIdx.x;
    PiState &st = A.states[j];
    const double s = D.sf ? D.sf[j] : 1.0;
    const CT *base = reinterpret_cast<const CT *>(A.C.ptr);
    // sharded run: D.G, A.nat and D.pi_pred span ALL genes and the cell's counts come from the gathered buffer
    const int32_t *gathered = A.tail_counts ? A.tail_counts + (size_t)blockIdx.x * A.tail_ld : nullptr;
    const bool hoist = M.drop_model == LPB200_DROP_LOGISTIC && D.P == 0;   // as pi_eval_kernel<.., HOIST = true>: same bits
    const double *xtab = A.xphi_tab;
    const int xw = A.xphi_w;
    for (;;) {
        if (tid == 0) {
            int req = st.req;
            int nk = (req == LP_REQ_F) ? 1 : (req == LP_REQ_GRAD ? 2 * q : 0);
            for (int k = 0; k < nk; k++) {
                double th[LP_PI_NP];
                for (int i = 0; i < q; i++) th[i] = st.b[i];
                if (req == LP_REQ_GRAD) {
                    int c = bfgs_grad_coord(k);
                    th[c] = bfgs_grad_point(st.b[c], k, A.ndeps);
                }
                lp_pi_transform(M.drop_model, q, th, s_coef[k]);
                if (hoist) {
                    s_pi[k] = lp_dropout_rate(1.0 * s_coef[k][0]);
                    s_l1m[k] = log(1.0 - s_pi[k]);
                }
            }
            s_nk = nk;
        }
        __syncthreads();
        const int nk = s_nk;
        if (nk == 0) break;
        // same summation structure as the lock-step path: 16-gene blocks summed in gene order, blocks
        // combined exactly (shared-memory integer atomics), so a cell finished here gets the bits it
        // would have got in lock-step rounds on any number of GPUs
        if (tid < LP_PI_KP) {
            s_ex[tid][0] = 0;
            s_ex[tid][1] = 0;
            s_ex[tid][2] = 0;
        }
        __syncthreads();
        const int n_blocks = (D.G + LP_PI_BLOCK - 1) / LP_PI_BLOCK;
        for (int blk = tid; blk < n_blocks; blk += blockDim.x) {
            double acc[LP_PI_KP];
#pragma unroll
            for (int k = 0; k < LP_PI_KP; k++) acc[k] = 0.0;
            const int i1 = min(D.G, (blk + 1) * LP_PI_BLOCK);
            for (int i = blk * LP_PI_BLOCK; i < i1; i++) {
                const int x = gathered ? gathered[i] : lp_load_count(base + (size_t)i * A.C.pitch, j);
                if (x < 0) continue;
                PointNat P;
                lp_nat_load(M, A.nat + (size_t)i * n_nat, P);
                const double mu = lp_mu_at(D, M, P, j);
                const double phi = lp_disp_at(D, M, P, j);
                const double mus = mu * s;
                const double lmu = (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) ? log(mu) : 0.0;
                double core;
                if (x == 0) core = exp(phi * log(phi / (phi + mus)));
                else if (lp_nb_small_x((double)x, phi)) core = lp_dnbinom_small_x((double)x, mus, phi);
                else core = ((xtab && x < xw) ? xtab[(size_t)i * xw + x] : lp_nb_xphi_part((double)x, phi)) + lp_nb_mu_part((double)x, mus, phi);
#pragma unroll
                for (int k = 0; k < LP_PI_KP; k++) {
                    if (k < nk) {
                        double v;
                        if (hoist) {
                            v = (x == 0) ? log((1.0 - s_pi[k]) * core + s_pi[k]) : s_l1m[k] + core;
                        } else {
                            int u = 1;
                            double lin = 1.0 * s_coef[k][0];
                            if (M.drop_model == LPB200_DROP_LOGISTIC_OFMU) lin += lmu * s_coef[k][u++];
                            for (int p = 0; p < D.P; p++, u++) lin += D.pi_pred[i + (size_t)p * D.G] * s_coef[k][u];
                            double pi = lp_dropout_rate(lin);
                            v = (x == 0) ? log((1.0 - pi) * core + pi) : log(1.0 - pi) + core;
                        }
                        if (v < LP_LL_FLOOR) v = LP_LL_FLOOR;
                        acc[k] += v;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < LP_PI_KP; k++) {
                if (k < nk) {
                    ExactSum e = {0, 0, 0};
                    lp_exact_add(e, acc[k]);
                    atomicAdd((unsigned long long *)&s_ex[k][0], (unsigned long long)e.a);
                    atomicAdd((unsigned long long *)&s_ex[k][1], (unsigned long long)e.b);
                    if (e.bad) atomicAdd((unsigned long long *)&s_ex[k][2], 1ull);
                }
            }
        }
        __syncthreads();
        if (tid < nk) s_vals[tid] = lp_exact_value(s_ex[tid][0], s_ex[tid][1], s_ex[tid][2]);
        __syncthreads();
        if (tid == 0) bfgs_advance(st, s_vals);
        __syncthreads();
    }
}

// sharded tail: this shard's counts of the remaining cells into the gathered cell-major buffer (columns
// [g_off, g_off + G) of every cell's row; the other shards' columns stay zero for the summing all-reduce)
template <typename CT>
__global__ void pi_gather_counts_kernel(CountView C, int G, const int *list, int n_tail, int g_off, int32_t *out, size_t ld) {
    const int k = blockIdx.y;
    if (k >= n_tail) return;
    const int j = list[k];
    const CT *base = reinterpret_cast<const CT *>(C.ptr);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < G; i += gridDim.x * blockDim.x)
        out[(size_t)k * ld + g_off + i] = lp_load_count(base + (size_t)i * C.pitch, j);
}

// ForAllCells: total[k] = sum_j vals[k][j]  (one CTA per point, fixed order)
__global__ void pi_total_kernel(const double *vals, int N, double *total) {
    __shared__ double sh[32];
    int k = blockIdx.x;
    double s = 0.0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) s += vals[(size_t)k * N + j];
    s = lp_block_sum(s, sh);
    if (threadIdx.x == 0) total[k] = s;
}

// ---------------------------------------------------------------------------------------------
// calcPostDrop_Matrix and the fit extraction of getFits.R
// ---------------------------------------------------------------------------------------------
template <typename CT>
__global__ void expand_kernel(DevDesign D, DevModel M, CountView C, const double *nat, const double *drop,
                              const int32_t *ids, int n_ids, int what, double *z) {
    __shared__ PointNat P;
    int r = blockIdx.y;
    int i = ids[r];
    if (threadIdx.x == 0) lp_nat_load(M, nat + (size_t)i * lp_n_nat(M), P);
    __syncthreads();
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D.N) return;
    const int kind = what & 15;
    double out;
    if (kind == LPB200_EXPAND_MEAN) {
        out = lp_mu_at(D, M, P, j);
    } else if (kind == LPB200_EXPAND_DISP) {
        out = lp_disp_at(D, M, P, j);
    } else if (kind == LPB200_EXPAND_DROP) {
        out = lp_dropout_rate(lp_drop_lin(D, M, drop, i, j, lp_mu_at(D, M, P, j)));
    } else {
        int x = lp_load_count(reinterpret_cast<const CT *>(C.ptr) + (size_t)i * C.pitch, j);
        if (kind == LPB200_EXPAND_NORMDATA) {  // getNormData (getFits.R:52-98)
            if (x < 0) out = NAN;
            else {
                out = (double)x;
                if (!(what & LPB200_EXPAND_NO_DEPTH)) out = out / (D.sf ? D.sf[j] : 1.0);
                if (!(what & LPB200_EXPAND_NO_BATCH)) {
                    double bp = 1.0;
                    int o = 0;
                    for (int c = 0; c < D.n_conf_mu; c++) {
                        bp = bp * P.bmu[o + D.bidx_mu[(size_t)c * D.N + j]];
                        o += D.nb_mu[c];
                    }
                    out = out / bp;
                }
            }
        } else if (x < 0) out = NAN;
        else if (x > 0) out = 0.0;
        else {
            double mu = lp_mu_at(D, M, P, j);   // no size factor (calcPostDrop.R:110)
            double phi = lp_disp_at(D, M, P, j);
            double pi = lp_dropout_rate(lp_drop_lin(D, M, drop, i, j, mu));
            out = lp_post_drop_zero(mu, phi, pi);
        }
    }
    z[r + (size_t)j * n_ids] = out;
}

// ---------------------------------------------------------------------------------------------
// count-matrix storage (K0): R column-major doubles / dgCMatrix -> gene-major integers
// ---------------------------------------------------------------------------------------------
// slab = columns [j0, j0+nj) of a G x N column-major double matrix; out gene-major int32
__global__ void pack_dense_kernel(const double *slab, int G, int j0, int nj, int32_t *out, size_t pitch) {
    __shared__ double tile[32][33];
    int gi = blockIdx.x * 32, jj = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: cell within tile, threadIdx.x: gene (contiguous)
        int g = gi + threadIdx.x, j = jj + r;
        tile[r][threadIdx.x] = (g < G && j < nj) ? slab[(size_t)j * G + g] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: gene within tile, threadIdx.x: cell (contiguous)
        int g = gi + r, j = jj + threadIdx.x;
        if (g < G && j < nj) {
            double v = tile[threadIdx.x][r];
            out[(size_t)g * pitch + j0 + j] = isnan(v) ? LPB200_NA_COUNT : (int32_t)llrint(v);
        }
    }
}

__global__ void scatter_csc_kernel(const int32_t *p, const int32_t *i, const double *x, int N, int32_t *out, size_t pitch) {
    int j = blockIdx.x;
    for (int e = p[j] + threadIdx.x; e < p[j + 1]; e += blockDim.x) {
        double v = x[e];
        out[(size_t)i[e] * pitch + j] = isnan(v) ? LPB200_NA_COUNT : (int32_t)llrint(v);
    }
}

// stable partition of one gene's cells: non-zero observations first, then zeros; NA dropped
__global__ void __launch_bounds__(256) build_perm_kernel(const int32_t *c, int G, int N, size_t pitch, uint32_t *perm,
                                                         int32_t *nnz_out, int32_t *nobs_out) {
    __shared__ int sh_nz[8], sh_z[8];
    __shared__ int run_nz, run_z, tot_nz;
    const int g = blockIdx.x;
    const int32_t *row = c + (size_t)g * pitch;
    uint32_t *out = perm + (size_t)g * pitch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cnt = 0;
    for (int j = tid; j < N; j += blockDim.x) cnt += row[j] > 0;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) sh_nz[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += sh_nz[w];
        tot_nz = t;
        run_nz = 0;
        run_z = 0;
    }
    __syncthreads();
    const int nnz = tot_nz;
    for (int base = 0; base < N; base += blockDim.x) {
        const int j = base + tid;
        const int v = (j < N) ? row[j] : -1;
        const unsigned bnz = __ballot_sync(0xffffffffu, v > 0), bz = __ballot_sync(0xffffffffu, v == 0);
        const unsigned lt = (1u << lane) - 1u;
        if (lane == 0) {
            sh_nz[warp] = __popc(bnz);
            sh_z[warp] = __popc(bz);
        }
        __syncthreads();
        int off_nz = run_nz, off_z = run_z;
        for (int w = 0; w < warp; w++) {
            off_nz += sh_nz[w];
            off_z += sh_z[w];
        }
        if (v > 0) out[off_nz + __popc(bnz & lt)] = (uint32_t)j;
        else if (v == 0) out[nnz + off_z + __popc(bz & lt)] = (uint32_t)j;
        __syncthreads();
        if (tid == 0) {
            for (int w = 0; w < 8; w++) {
                run_nz += sh_nz[w];
                run_z += sh_z[w];
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        nnz_out[g] = nnz;
        nobs_out[g] = nnz + run_z;
    }
}

// per gene: sum, number of non-NA cells, max
__global__ void row_stats_kernel(const int32_t *c, int G, int N, size_t pitch, double *mean, int32_t *xmax) {
    __shared__ double sh[32];
    __shared__ int shi[32];
    int g = blockIdx.x;
    const int32_t *row = c + (size_t)g * pitch;
    double s = 0.0, n = 0.0;
    int mx = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        int v = row[j];
        if (v >= 0) {
            s += (double)v;
            n += 1.0;
            mx = max(mx, v);
        }
    }
    s = lp_block_sum(s, sh);
    n = lp_block_sum(n, sh);
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) shi[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) mx = max(mx, shi[w]);
        mean[g] = s / n;
        xmax[g] = mx;
    }
}

// per cell: largest count over this shard's genes (NA skipped); per launch: number of negative entries (the NA marker -1
// included) -- the input checks and the all-zero cell filter of processSCData.R:102-112,276-289 on the uploaded matrix
__global__ void col_stats_kernel(const int32_t *c, int G, int N, size_t pitch, int32_t *colmax, unsigned long long *n_negative) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    int mx = 0;
    unsigned long long neg = 0;
    for (int g = 0; g < G; g++) {
        const int v = c[(size_t)g * pitch + j];
        mx = max(mx, v);
        neg += (v < 0);
    }
    colmax[j] = mx;
    if (neg) atomicAdd(n_negative, neg);
}

__global__ void narrow_u16_kernel(const int32_t *in, uint16_t *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int v = in[i];
        out[i] = v < 0 ? (uint16_t)LP_NA_U16 : (uint16_t)v;
    }
}

// splines initial values: coef = Pinv y, y = log(x + 1)  (initialiseMuModel.R:111-118)
template <typename CT>
__global__ void __launch_bounds__(256) spline_init_kernel(DevDesign D, CountView C, const double *Pinv /* K x N row-major */,
                                                          int K, double *coef /* G x K col-major */) {
    __shared__ double sh[32];
    for (int gene = blockIdx.x; gene < D.G; gene += gridDim.x) {
        const CT *row = reinterpret_cast<const CT *>(C.ptr) + (size_t)gene * C.pitch;
        for (int k = 0; k < K; k++) {
            double s = 0.0;
            for (int j = threadIdx.x; j < D.N; j += blockDim.x)
                s += Pinv[(size_t)k * D.N + j] * log((double)max(lp_load_count(row, j), 0) + 1.0);   // NA as 0
            s = lp_block_sum(s, sh);
            if (threadIdx.x == 0) coef[gene + (size_t)k * D.G] = s;
        }
    }
}

// FP64 peak probe: 8 independent DFMA chains per thread
__global__ void fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// exp / log of the inner loop (lp_libm.cuh) next to the CUDA math library's, for the bit-identity test
__global__ void libm_probe_kernel(const double *x, int n, int use_library, double *e, double *l) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        if (use_library) {
            e[k] = exp(x[k]);
            l[k] = log(x[k]);
        } else {
            // even entries through the single routine, odd entries through the pair routine (partner: the
            // previous argument), so both are compared with the library
            double partner;
            if (k & 1) lp_exp_pair(x[k - 1], x[k], partner, e[k]);
            else e[k] = lp_exp(x[k]);
            // positive normal finite arguments also go through the routine without the special-case tests
            const bool plain = x[k] >= 2.2250738585072014e-308 && x[k] <= 1.7976931348623157e308;
            l[k] = (plain && (k & 2)) ? lp_logp<true>(x[k]) : lp_log(x[k]);
        }
    }
}

// a / b and 1 / b by the check-free fast paths of lp_libm.cuh next to the compiler's division
__global__ void div_probe_kernel(const double *a, const double *b, int n, int use_library, double *q, double *r) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        if (use_library) {
            q[k] = a[k] / b[k];
            r[k] = 1.0 / b[k];
        } else {
            q[k] = lp_div<true>(a[k], b[k]);
            r[k] = lp_rcp<true>(b[k]);
        }
    }
}

// MUFU.RCP64H (rcp.approx.ftz.f64) of the double (hi, lo): upper and lower result words.  The logarithm of
// lp_libm.cuh is the only routine whose bits depend on this seed (the divisions are correctly rounded); the
// dump of all 2^20 mantissas (scripts/dump_rcp64h.py -> tests/golden/rcp64h.npz) lets the host mirror of the
// fit kernel reproduce the device logarithm bit for bit.
__global__ void rcp64h_probe_kernel(const int32_t *hi, int n, int32_t lo, int32_t *out_hi, int32_t *out_lo) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double r = lp_rcp_approx(__hiloint2double(hi[k], lo));
        out_hi[k] = __double2hiint(r);
        out_lo[k] = __double2loint(r);
    }
}
