/*
 * rapidnet_oracle.c -- CPU restatement of RapidNet's APG stochastic-MPC path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path and the
 * "port" CPU baseline of bench.py; nothing in the product imports, links or
 * executes it.  It restates, operation by operation and in fp32, what the
 * reference does on the GPU through cuBLAS (all matrices column-major):
 *
 *   factor step            /root/reference/src/Engine.cu:382-463, 671-774, 1318-1359
 *   preconditioning kernels /root/reference/src/Utilities.cu:33-58, 360-405
 *   per-solve affine terms /root/reference/src/Engine.cu:1147-1316,
 *                          /root/reference/src/Utilities.cu:69-131
 *   APG steps              /root/reference/src/SmpcController.cu:420-450, 535-864
 *   tree kernels           /root/reference/src/Utilities.cu:142-201
 *   prox kernels           /root/reference/src/Utilities.cu:237-304
 *   loop, theta, infeas.   /root/reference/src/SmpcController.cu:1480-1525
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file against every
 * golden vector the reference's own tests hold for the path
 * (src/test/testDataFiles/engineTest.json, smpcTest.json -> tests/golden/toy.npz).
 *
 * Third-party arithmetic not in /root/reference: cuBLAS 12.9 (gemm/gemv/axpy/
 * nrm2/isamax/getrf/getri batched) and cuSOLVER Dgesvd.  Their published
 * semantics are restated with plain loops; the null-space basis (SVD) is an
 * input (orc_set_L) or computed by Householder QR (orc_null_space) -- any
 * orthonormal basis of null(E) yields the same u, x, xi, psi.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* real = float: the oracle proper (fp32 like the reference).  -DORC_F64 builds the same code in double
 * (librapidnet_oracle64.so): the rounding-free trajectory that tests use to measure the fp32 noise floor of
 * the reference itself (tests/test_gpu_reference.py, DESIGN.md "tolerances"). */
#ifdef ORC_F64
typedef double real;
#define r_sqrt sqrt
#define r_abs fabs
#else
typedef float real;
#define r_sqrt sqrtf
#define r_abs fabsf
#endif
typedef float f32;

typedef struct orc {
    int nx, nu, nd, ne, nv, N, K, nodes, n_nonleaf, fb;
    /* tree (copies) */
    int *stages, *nps, *cum, *ancestor, *nchild, *nchild_cum;
    real *prob, *err_demand, *err_price;
    /* network */
    real *B, *Gd, *E, *Ed, *xmin, *xmax, *xsafe, *umin, *umax, *alpha1;
    /* config */
    real *W, *precond, pen_x, pen_xs, step, w_econ;
    /* factor-step outputs (reference layouts) */
    real *L, *Lhat;                 /* nu*nv, nu*nd */
    real *Wv;                       /* nu*nv  = W L        (devMatWv) */
    real *Rbar;                     /* nv*nv  = L' W L     */
    real *Bbar;                     /* nv*nx  = (B L)'     (= G) */
    real *s_u, *s_x, *s_xs;         /* per node diagonals of sysG / sysF */
    real *sxmin, *sxmax, *sxs, *sumin, *sumax; /* scaled bounds per node */
    real *Omega, *Theta;            /* fbn * nv*nv, fbn * nv*nx */
    int fbn;                         /* number of distinct Omega/Theta */
    real *Phi, *Psi, *D, *F;        /* per node */
    /* per-solve */
    real *xcur, *uprev, *dprev, *uhat_prev;
    real *e, *uhat, *alpha, *beta;
    /* APG state */
    real *X, *U, *V, *sigma;
    real *xi, *psi;                 /* y_{k-1} */
    real *upd_xi, *upd_psi;         /* y_k */
    real *acc_xi, *acc_psi;         /* w */
    real *pri_xi, *pri_psi;         /* Hx */
    real *dual_xi, *dual_psi;       /* z */
    real *res_xi, *res_psi;
    real *Q, *R;                    /* K*nx, K*nv */
    real dist_x, dist_xs;           /* last prox distances */
    int have_L;
} orc;

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: out of memory (%zu x %zu)\n", n, sz); abort(); }
    return p;
}
static real *fdup(const real *src, size_t n) {
    real *p = (real *)xcalloc(n, sizeof(real));
    if (src) memcpy(p, src, n * sizeof(real));
    return p;
}
static int *idup(const int *src, size_t n) {
    int *p = (int *)xcalloc(n, sizeof(int));
    if (src) memcpy(p, src, n * sizeof(int));
    return p;
}

/* y(m) = alpha*A(m x n, ld)*x + beta_*y ; column-major, column-sweep accumulation */
static void gemv_n(int m, int n, real alpha, const real *A, int ld, const real *x, real beta_, real *y) {
    if (beta_ == 0.0f) { for (int r = 0; r < m; r++) y[r] = 0.0f; }
    else if (beta_ != 1.0f) { for (int r = 0; r < m; r++) y[r] *= beta_; }
    for (int j = 0; j < n; j++) {
        const real xj = alpha * x[j];
        const real *a = A + (size_t)j * ld;
        for (int r = 0; r < m; r++) y[r] += a[r] * xj;
    }
}
/* y(n) = alpha*A'(A is m x n) * x(m) + beta_*y */
static void gemv_t(int m, int n, real alpha, const real *A, int ld, const real *x, real beta_, real *y) {
    for (int j = 0; j < n; j++) {
        const real *a = A + (size_t)j * ld;
        real s = 0.0f;
        for (int r = 0; r < m; r++) s += a[r] * x[r];
        y[j] = alpha * s + (beta_ == 0.0f ? 0.0f : beta_ * y[j]);
    }
}
/* C(m x n) = alpha * op(A) * op(B) + beta_*C ; generic small gemm (setup only) */
static void gemm(int ta, int tb, int m, int n, int k, real alpha, const real *A, int lda,
                 const real *B, int ldb, real beta_, real *C, int ldc) {
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) {
            real s = 0.0f;
            for (int p = 0; p < k; p++) {
                real a = ta ? A[p + (size_t)i * lda] : A[i + (size_t)p * lda];
                real b = tb ? B[j + (size_t)p * ldb] : B[p + (size_t)j * ldb];
                s += a * b;
            }
            C[i + (size_t)j * ldc] = alpha * s + (beta_ == 0.0f ? 0.0f : beta_ * C[i + (size_t)j * ldc]);
        }
}

/* fp32 LU with partial pivoting + inverse (cublasSgetrfBatched/SgetriBatched semantics). returns 0 ok */
static int inverse_lu(int n, real *A /* destroyed */, real *inv) {
    int *piv = (int *)xcalloc(n, sizeof(int));
    for (int k = 0; k < n; k++) {
        int p = k; real mx = r_abs(A[k + (size_t)k * n]);
        for (int i = k + 1; i < n; i++) { real v = r_abs(A[i + (size_t)k * n]); if (v > mx) { mx = v; p = i; } }
        piv[k] = p;
        if (mx == 0.0f) { free(piv); return k + 1; }
        if (p != k) for (int j = 0; j < n; j++) { real t = A[k + (size_t)j * n]; A[k + (size_t)j * n] = A[p + (size_t)j * n]; A[p + (size_t)j * n] = t; }
        real d = 1.0f / A[k + (size_t)k * n];
        for (int i = k + 1; i < n; i++) A[i + (size_t)k * n] *= d;
        for (int j = k + 1; j < n; j++) {
            real akj = A[k + (size_t)j * n];
            for (int i = k + 1; i < n; i++) A[i + (size_t)j * n] -= A[i + (size_t)k * n] * akj;
        }
    }
    /* solve A X = I column by column: P A = L U */
    for (int c = 0; c < n; c++) {
        real *x = inv + (size_t)c * n;
        for (int i = 0; i < n; i++) x[i] = (i == c) ? 1.0f : 0.0f;
        for (int k = 0; k < n; k++) if (piv[k] != k) { real t = x[k]; x[k] = x[piv[k]]; x[piv[k]] = t; }
        for (int k = 0; k < n; k++) { real xk = x[k]; if (xk != 0.0f) for (int i = k + 1; i < n; i++) x[i] -= A[i + (size_t)k * n] * xk; }
        for (int k = n - 1; k >= 0; k--) { x[k] /= A[k + (size_t)k * n]; real xk = x[k]; for (int i = 0; i < k; i++) x[i] -= A[i + (size_t)k * n] * xk; }
    }
    free(piv);
    return 0;
}

orc *orc_create(int nx, int nu, int nd, int ne, int nv, int N, int K, int nodes, int n_nonleaf,
                const int *stages, const int *nps /*N+1*/, const int *cum /*N+2*/, const int *ancestor,
                const int *nchild, const int *nchild_cum, const real *prob,
                const real *err_demand, const real *err_price,
                const real *B, const real *Gd, const real *E, const real *Ed,
                const real *xmin, const real *xmax, const real *xsafe, const real *umin,
                const real *umax, const real *alpha1,
                const real *W, const real *precond, real pen_x, real pen_xs, real step) {
    orc *o = (orc *)xcalloc(1, sizeof(orc));
    o->nx = nx; o->nu = nu; o->nd = nd; o->ne = ne; o->nv = nv; o->N = N; o->K = K; o->nodes = nodes;
    o->n_nonleaf = n_nonleaf;
    o->stages = idup(stages, nodes); o->nps = idup(nps, N + 1); o->cum = idup(cum, N + 2);
    o->ancestor = idup(ancestor, nodes); o->nchild = idup(nchild, n_nonleaf);
    o->nchild_cum = idup(nchild_cum, nodes); o->prob = fdup(prob, nodes);
    o->err_demand = fdup(err_demand, (size_t)nodes * nd); o->err_price = fdup(err_price, (size_t)nodes * nu);
    o->B = fdup(B, (size_t)nx * nu); o->Gd = fdup(Gd, (size_t)nx * nd);
    o->E = fdup(E, (size_t)ne * nu); o->Ed = fdup(Ed, (size_t)ne * nd);
    o->xmin = fdup(xmin, nx); o->xmax = fdup(xmax, nx); o->xsafe = fdup(xsafe, nx);
    o->umin = fdup(umin, nu); o->umax = fdup(umax, nu); o->alpha1 = fdup(alpha1, nu);
    o->W = fdup(W, (size_t)nu * nu); o->precond = fdup(precond, (size_t)N * (nu + 2 * nx));
    o->pen_x = pen_x; o->pen_xs = pen_xs; o->step = step; o->w_econ = 1.0f; /* SmpcConfiguration.cu:38-40 */
    /* ScenarioTree.cu:149-158 */
    o->fb = 0;
    for (int s = 0; s < N - 1; s++) if (nps[s] == nps[s + 1]) { o->fb = cum[s + 1]; break; }
    o->fbn = o->fb > 0 ? o->fb : nodes;
    size_t n = (size_t)nodes;
    o->L = fdup(NULL, (size_t)nu * nv); o->Lhat = fdup(NULL, (size_t)nu * nd);
    o->Wv = fdup(NULL, (size_t)nu * nv); o->Rbar = fdup(NULL, (size_t)nv * nv); o->Bbar = fdup(NULL, (size_t)nv * nx);
    o->s_u = fdup(NULL, n * nu); o->s_x = fdup(NULL, n * nx); o->s_xs = fdup(NULL, n * nx);
    o->sxmin = fdup(NULL, n * nx); o->sxmax = fdup(NULL, n * nx); o->sxs = fdup(NULL, n * nx);
    o->sumin = fdup(NULL, n * nu); o->sumax = fdup(NULL, n * nu);
    o->Omega = fdup(NULL, (size_t)o->fbn * nv * nv); o->Theta = fdup(NULL, (size_t)o->fbn * nv * nx);
    o->Phi = fdup(NULL, n * 2 * nv * nx); o->D = fdup(NULL, n * 2 * nv * nx);
    o->Psi = fdup(NULL, n * nv * nu); o->F = fdup(NULL, n * nv * nu);
    o->xcur = fdup(NULL, nx); o->uprev = fdup(NULL, nu); o->dprev = fdup(NULL, nd); o->uhat_prev = fdup(NULL, nu);
    o->e = fdup(NULL, n * nx); o->uhat = fdup(NULL, n * nu); o->alpha = fdup(NULL, n * nu); o->beta = fdup(NULL, n * nv);
    o->X = fdup(NULL, n * nx); o->U = fdup(NULL, n * nu); o->V = fdup(NULL, n * nv); o->sigma = fdup(NULL, n * nv);
    o->xi = fdup(NULL, n * 2 * nx); o->psi = fdup(NULL, n * nu);
    o->upd_xi = fdup(NULL, n * 2 * nx); o->upd_psi = fdup(NULL, n * nu);
    o->acc_xi = fdup(NULL, n * 2 * nx); o->acc_psi = fdup(NULL, n * nu);
    o->pri_xi = fdup(NULL, n * 2 * nx); o->pri_psi = fdup(NULL, n * nu);
    o->dual_xi = fdup(NULL, n * 2 * nx); o->dual_psi = fdup(NULL, n * nu);
    o->res_xi = fdup(NULL, n * 2 * nx); o->res_psi = fdup(NULL, n * nu);
    /* widest stage bounds the scratch (the reference sizes it with K) */
    int widest = 1; for (int s = 0; s < N; s++) if (nps[s] > widest) widest = nps[s];
    o->Q = fdup(NULL, (size_t)widest * nx); o->R = fdup(NULL, (size_t)widest * nv);
    return o;
}

void orc_destroy(orc *o) {
    if (!o) return;
    void *ptrs[] = {o->stages, o->nps, o->cum, o->ancestor, o->nchild, o->nchild_cum, o->prob, o->err_demand,
        o->err_price, o->B, o->Gd, o->E, o->Ed, o->xmin, o->xmax, o->xsafe, o->umin, o->umax, o->alpha1, o->W,
        o->precond, o->L, o->Lhat, o->Wv, o->Rbar, o->Bbar, o->s_u, o->s_x, o->s_xs, o->sxmin, o->sxmax, o->sxs,
        o->sumin, o->sumax, o->Omega, o->Theta, o->Phi, o->Psi, o->D, o->F, o->xcur, o->uprev, o->dprev,
        o->uhat_prev, o->e, o->uhat, o->alpha, o->beta, o->X, o->U, o->V, o->sigma, o->xi, o->psi, o->upd_xi,
        o->upd_psi, o->acc_xi, o->acc_psi, o->pri_xi, o->pri_psi, o->dual_xi, o->dual_psi, o->res_xi, o->res_psi,
        o->Q, o->R};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
    free(o);
}

void orc_set_L(orc *o, const real *L, const real *Lhat) {
    memcpy(o->L, L, (size_t)o->nu * o->nv * sizeof(real));
    memcpy(o->Lhat, Lhat, (size_t)o->nu * o->nd * sizeof(real));
    o->have_L = 1;
}

/*
 * L = orthonormal basis of null(E), Lhat = -pinv(E) Ed (Engine.cu:466-669 uses a
 * double-precision SVD of E'; here: Householder QR of E' in double, E' = Q R,
 * L = Q(:, ne:nu), pinv(E) = Q1 R^-T).  Spans the same subspace; the basis
 * differs from cuSOLVER's by an orthogonal rotation.
 */
int orc_null_space(orc *o) {
    const int nu = o->nu, ne = o->ne, nd = o->nd, nv = o->nv;
    if (nv != nu - ne) return 1;
    double *A = (double *)xcalloc((size_t)nu * ne, sizeof(double));   /* E' : nu x ne */
    double *Q = (double *)xcalloc((size_t)nu * nu, sizeof(double));
    double *v = (double *)xcalloc(nu, sizeof(double));
    for (int r = 0; r < ne; r++) for (int c = 0; c < nu; c++) A[c + (size_t)r * nu] = o->E[r + (size_t)c * ne];
    for (int i = 0; i < nu; i++) Q[i + (size_t)i * nu] = 1.0;
    for (int k = 0; k < ne; k++) {
        double nrm = 0; for (int i = k; i < nu; i++) nrm += A[i + (size_t)k * nu] * A[i + (size_t)k * nu];
        nrm = sqrt(nrm);
        if (nrm == 0.0) { free(A); free(Q); free(v); return 2; }
        double a0 = A[k + (size_t)k * nu]; double alpha = a0 >= 0 ? -nrm : nrm;
        for (int i = 0; i < nu; i++) v[i] = 0; for (int i = k; i < nu; i++) v[i] = A[i + (size_t)k * nu];
        v[k] -= alpha;
        double vn = 0; for (int i = k; i < nu; i++) vn += v[i] * v[i];
        if (vn == 0.0) continue;
        for (int j = 0; j < ne; j++) { double s = 0; for (int i = k; i < nu; i++) s += v[i] * A[i + (size_t)j * nu]; s = 2 * s / vn; for (int i = k; i < nu; i++) A[i + (size_t)j * nu] -= s * v[i]; }
        for (int j = 0; j < nu; j++) { /* Q = Q H */
            double s = 0; for (int i = k; i < nu; i++) s += Q[j + (size_t)i * nu] * v[i]; s = 2 * s / vn;
            for (int i = k; i < nu; i++) Q[j + (size_t)i * nu] -= s * v[i];
        }
    }
    for (int c = 0; c < nv; c++) for (int r = 0; r < nu; r++) o->L[r + (size_t)c * nu] = (real)Q[r + (size_t)(ne + c) * nu];
    /* Lhat = -Q1 R^-T Ed : solve R' Y = Ed (ne x nd), then Q1 Y */
    double *Y = (double *)xcalloc((size_t)ne * nd, sizeof(double));
    for (int c = 0; c < nd; c++) {
        for (int i = 0; i < ne; i++) {
            double s = o->Ed[i + (size_t)c * ne];
            for (int p = 0; p < i; p++) s -= A[p + (size_t)i * nu] * Y[p + (size_t)c * ne]; /* R'(i,p) = R(p,i) */
            Y[i + (size_t)c * ne] = s / A[i + (size_t)i * nu];
        }
        for (int r = 0; r < nu; r++) { double s = 0; for (int p = 0; p < ne; p++) s += Q[r + (size_t)p * nu] * Y[p + (size_t)c * ne]; o->Lhat[r + (size_t)c * nu] = (real)(-s); }
    }
    free(A); free(Q); free(v); free(Y);
    o->have_L = 1;
    return 0;
}

/* Engine::initialiseSystemDevice + Engine::factorStep */
int orc_factor_step(orc *o) {
    const int nx = o->nx, nu = o->nu, nv = o->nv, N = o->N, nodes = o->nodes, K = o->K;
    if (!o->have_L) { int rc = orc_null_space(o); if (rc) return 100 + rc; }
    /* Wv = W L ; Rbar = L' Wv   (Engine.cu:412-416) */
    gemm(0, 0, nu, nv, nu, 1.0f, o->W, nu, o->L, nu, 0.0f, o->Wv, nu);
    gemm(1, 0, nv, nv, nu, 1.0f, o->L, nu, o->Wv, nu, 0.0f, o->Rbar, nv);
    /* bounds and preconditioning (Engine.cu:421-451, Utilities.cu:33-58, 360-405) */
    for (int s = 0; s < N; s++) {
        const real *pc = o->precond + (size_t)s * (2 * nx + nu);
        for (int j = 0; j < o->nps[s]; j++) {
            int i = o->cum[s] + j;
            real sp = r_sqrt(o->prob[i]);
            for (int t = 0; t < nu; t++) {
                real sc = sp * pc[t];
                o->s_u[(size_t)i * nu + t] = sc;
                o->sumax[(size_t)i * nu + t] = sc * o->umax[t];
                o->sumin[(size_t)i * nu + t] = sc * o->umin[t];
            }
            for (int t = 0; t < nx; t++) {
                real scx = sp * pc[nu + t], scs = sp * pc[nu + nx + t];
                o->s_x[(size_t)i * nx + t] = scx; o->s_xs[(size_t)i * nx + t] = scs;
                o->sxmax[(size_t)i * nx + t] = scx * o->xmax[t];
                o->sxmin[(size_t)i * nx + t] = scx * o->xmin[t];
                o->sxs[(size_t)i * nx + t] = scs * o->xsafe[t];
            }
        }
    }
    /* Bbar' = L' B'  (Engine.cu:702) */
    gemm(1, 1, nv, nx, nu, 1.0f, o->L, nu, o->B, nx, 0.0f, o->Bbar, nv);
    /* Omega_i = (p_i Rbar)^-1 for the nodes of stages that start before fb (Engine.cu:707-714) */
    int bad = 0;
    #pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < o->fbn; i++) {
        real *tmp = (real *)xcalloc((size_t)nv * nv, sizeof(real));
        for (int t = 0; t < nv * nv; t++) tmp[t] = o->Rbar[t] * o->prob[i];   /* Sscal by p (Engine.cu:434) */
        if (inverse_lu(nv, tmp, o->Omega + (size_t)i * nv * nv)) bad = 1;
        free(tmp);
    }
    if (bad) return 3;
    #pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < o->fbn; i++)   /* Theta = -0.5 Omega Bbar' (Engine.cu:735-737) */
        gemm(0, 0, nv, nx, nv, -0.5f, o->Omega + (size_t)i * nv * nv, nv, o->Bbar, nv, 0.0f,
             o->Theta + (size_t)i * nv * nx, nv);
    /* per node F, D, Phi, Psi (Engine.cu:716-747) */
    #pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < nodes; i++) {
        int s = o->stages[i], rel = i - o->cum[s];
        int oi = (o->fb > 0 && o->fb <= o->cum[s]) ? o->fb - K + rel : i;   /* Engine.cu:210-221 */
        const real *Om = o->Omega + (size_t)oi * nv * nv;
        real *Fi = o->F + (size_t)i * nv * nu, *Di = o->D + (size_t)i * 2 * nv * nx;
        for (int c = 0; c < nu; c++) for (int r = 0; r < nv; r++)
            Fi[r + (size_t)c * nv] = o->L[c + (size_t)r * nu] * o->s_u[(size_t)i * nu + c];
        for (int c = 0; c < nx; c++) for (int r = 0; r < nv; r++) {
            Di[r + (size_t)c * nv] = o->Bbar[r + (size_t)c * nv] * o->s_x[(size_t)i * nx + c];
            Di[r + (size_t)(nx + c) * nv] = o->Bbar[r + (size_t)c * nv] * o->s_xs[(size_t)i * nx + c];
        }
        gemm(0, 0, nv, 2 * nx, nv, -0.5f, Om, nv, Di, nv, 0.0f, o->Phi + (size_t)i * 2 * nv * nx, nv);
        gemm(0, 0, nv, nu, nv, -0.5f, Om, nv, Fi, nv, 0.0f, o->Psi + (size_t)i * nv * nu, nv);
    }
    return 0;
}

/* Engine::updateStateControl (Engine.cu:1300-1316) */
void orc_update_state(orc *o, const real *x, const real *uprev, const real *dprev) {
    memcpy(o->xcur, x, o->nx * sizeof(real));
    memcpy(o->uprev, uprev, o->nu * sizeof(real));
    memcpy(o->dprev, dprev, o->nd * sizeof(real));
    gemv_n(o->nu, o->nd, 1.0f, o->Lhat, o->nu, o->dprev, 0.0f, o->uhat_prev);
}

/* Engine::eliminateInputDistubanceCoupling (Engine.cu:1147-1298) */
void orc_eliminate(orc *o, const real *dhat /*N*nd*/, const real *alphahat /*N*nu*/,
                   int demand_uncertainty, int price_uncertainty) {
    const int nx = o->nx, nu = o->nu, nv = o->nv, nd = o->nd, N = o->N, nodes = o->nodes;
    real *ah = fdup(alphahat, (size_t)N * nu);
    for (int s = 0; s < N; s++) for (int t = 0; t < nu; t++) ah[(size_t)s * nu + t] += o->alpha1[t];
    real *dU = fdup(NULL, (size_t)nodes * nu), *zeta = fdup(NULL, (size_t)nodes * nu);
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nodes; i++) {
        int s = o->stages[i];
        real *d = (real *)xcalloc(nd, sizeof(real));
        for (int t = 0; t < nd; t++) d[t] = (demand_uncertainty ? o->err_demand[(size_t)i * nd + t] : 0.0f) + dhat[(size_t)s * nd + t];
        gemv_n(nx, nd, 1.0f, o->Gd, nx, d, 0.0f, o->e + (size_t)i * nx);
        gemv_n(nu, nd, 1.0f, o->Lhat, nu, d, 0.0f, o->uhat + (size_t)i * nu);
        for (int t = 0; t < nu; t++) {
            real a = (price_uncertainty ? o->err_price[(size_t)i * nu + t] : 0.0f) + ah[(size_t)s * nu + t];
            o->alpha[(size_t)i * nu + t] = o->w_econ * a;
        }
        free(d);
    }
    /* calculateDiffUhat / calculateZeta (Utilities.cu:69-131) */
    for (int i = 0; i < nodes; i++) for (int t = 0; t < nu; t++) {
        if (i == 0) dU[t] = o->uhat[t] - o->uhat_prev[t];
        else dU[(size_t)i * nu + t] = o->uhat[(size_t)i * nu + t] - o->uhat[(size_t)(o->ancestor[i] - 1) * nu + t];
    }
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nodes; i++) {
        real *ab = (real *)xcalloc(nv, sizeof(real));
        for (int t = 0; t < nu; t++) {
            real z = o->prob[i] * dU[(size_t)i * nu + t];
            if (i < o->n_nonleaf) {
                int c0 = (i == 0) ? 1 : o->nchild_cum[i - 1] + 1;
                int nc = (i == 0) ? o->nchild_cum[0] : o->nchild_cum[i] - o->nchild_cum[i - 1];
                for (int c = 0; c < nc; c++) z = z - o->prob[c0 + c] * dU[(size_t)(c0 + c) * nu + t];
            }
            zeta[(size_t)i * nu + t] = z;
        }
        /* beta = 2 Wv' zeta ; beta += p * (L' alpha)   (Engine.cu:1245-1261) */
        gemv_t(nu, nv, 2.0f, o->Wv, nu, zeta + (size_t)i * nu, 0.0f, o->beta + (size_t)i * nv);
        gemv_t(nu, nv, 1.0f, o->L, nu, o->alpha + (size_t)i * nu, 0.0f, ab);
        for (int t = 0; t < nv; t++) o->beta[(size_t)i * nv + t] += o->prob[i] * ab[t];
        free(ab);
    }
    free(ah); free(dU); free(zeta);
}

/* SmpcController::initialiseAlgorithm (SmpcController.cu:420-450) */
void orc_apg_init(orc *o) {
    size_t n = (size_t)o->nodes;
    memset(o->xi, 0, n * 2 * o->nx * sizeof(real)); memset(o->psi, 0, n * o->nu * sizeof(real));
    memset(o->acc_xi, 0, n * 2 * o->nx * sizeof(real)); memset(o->acc_psi, 0, n * o->nu * sizeof(real));
    memset(o->pri_xi, 0, n * 2 * o->nx * sizeof(real)); memset(o->pri_psi, 0, n * o->nu * sizeof(real));
    memset(o->dual_xi, 0, n * 2 * o->nx * sizeof(real)); memset(o->dual_psi, 0, n * o->nu * sizeof(real));
    memset(o->upd_xi, 0, n * 2 * o->nx * sizeof(real)); memset(o->upd_psi, 0, n * o->nu * sizeof(real));
}

/* SmpcController::dualExtrapolationStep (SmpcController.cu:535-557) */
void orc_extrapolate(orc *o, real lambda) {
    size_t nxi = (size_t)o->nodes * 2 * o->nx, nps = (size_t)o->nodes * o->nu;
    real a1 = 1 + lambda, a2 = -lambda;
    #pragma omp parallel for schedule(static)
    for (size_t t = 0; t < nxi; t++) { real w = o->upd_xi[t] * a1; w += a2 * o->xi[t]; o->acc_xi[t] = w; o->xi[t] = o->upd_xi[t]; }
    #pragma omp parallel for schedule(static)
    for (size_t t = 0; t < nps; t++) { real w = o->upd_psi[t] * a1; w += a2 * o->psi[t]; o->acc_psi[t] = w; o->psi[t] = o->upd_psi[t]; }
}

/* SmpcController::solveStep (SmpcController.cu:563-755) */
void orc_solve_step(orc *o) {
    const int nx = o->nx, nu = o->nu, nv = o->nv, N = o->N, nodes = o->nodes, K = o->K;
    memcpy(o->sigma, o->beta, (size_t)nodes * nv * sizeof(real));
    int widest = 1; for (int s = 0; s < N; s++) if (o->nps[s] > widest) widest = o->nps[s];
    real *tq = fdup(NULL, (size_t)widest * nx), *tr = fdup(NULL, (size_t)widest * nv);
    for (int s = N - 1; s >= 0; s--) {
        const int c0 = o->cum[s], ns = o->nps[s];
        #pragma omp parallel for schedule(static)
        for (int j = 0; j < ns; j++) {
            const int i = c0 + j;
            const int oi = (o->fb > 0 && o->fb <= c0) ? o->fb - K + j : i;
            real *sg = o->sigma + (size_t)i * nv, *v = o->V + (size_t)i * nv;
            real *r = o->R + (size_t)j * nv, *q = o->Q + (size_t)j * nx;
            const real *wxi = o->acc_xi + (size_t)i * 2 * nx, *wps = o->acc_psi + (size_t)i * nu;
            if (s < N - 1) for (int t = 0; t < nv; t++) sg[t] += r[t];
            gemv_n(nv, nv, -0.5f, o->Omega + (size_t)oi * nv * nv, nv, sg, 0.0f, v);
            if (s < N - 1) gemv_n(nv, nx, 1.0f, o->Theta + (size_t)oi * nv * nx, nv, q, 1.0f, v);
            gemv_n(nv, nu, 1.0f, o->Psi + (size_t)i * nv * nu, nv, wps, 1.0f, v);
            gemv_n(nv, 2 * nx, 1.0f, o->Phi + (size_t)i * 2 * nv * nx, nv, wxi, 1.0f, v);
            memcpy(r, sg, nv * sizeof(real));
            gemv_n(nv, 2 * nx, 1.0f, o->D + (size_t)i * 2 * nv * nx, nv, wxi, 1.0f, r);
            gemv_n(nv, nu, 1.0f, o->F + (size_t)i * nv * nu, nv, wps, 1.0f, r);
            if (s < N - 1) gemv_n(nv, nx, 1.0f, o->Bbar, nv, q, 1.0f, r);
            /* q = sysF' xi (+ q): sysF = [diag(s_x); diag(s_xs)] */
            for (int t = 0; t < nx; t++) {
                real acc = o->s_x[(size_t)i * nx + t] * wxi[t] + o->s_xs[(size_t)i * nx + t] * wxi[nx + t];
                q[t] = (s < N - 1) ? acc + q[t] : acc;
            }
        }
        if (s > 0) {
            const int pn = o->nps[s - 1], pc0 = o->cum[s - 1];
            if (ns - pn > 0) {   /* solveSumChildren (Utilities.cu:168-201) */
                for (int p = 0; p < pn; p++) {
                    int node = pc0 + p;
                    int first = (node == 0 ? 0 : o->nchild_cum[node - 1]) - (pc0 == 0 ? 0 : o->nchild_cum[pc0 - 1]);
                    int nc = o->nchild[node];
                    for (int t = 0; t < nx; t++) { real a = o->Q[(size_t)first * nx + t]; for (int c = 1; c < nc; c++) a += o->Q[(size_t)(first + c) * nx + t]; tq[(size_t)p * nx + t] = a; }
                    for (int t = 0; t < nv; t++) { real a = o->R[(size_t)first * nv + t]; for (int c = 1; c < nc; c++) a += o->R[(size_t)(first + c) * nv + t]; tr[(size_t)p * nv + t] = a; }
                }
                memcpy(o->R, tr, (size_t)pn * nv * sizeof(real));
                memcpy(o->Q, tq, (size_t)pn * nx * sizeof(real));
            }
        }
    }
    /* forward substitution (SmpcController.cu:675-741) */
    memcpy(o->U, o->uhat, (size_t)nodes * nu * sizeof(real));
    for (int s = 0; s < N; s++) {
        const int c0 = o->cum[s], ns = o->nps[s];
        if (s == 0) {
            for (int t = 0; t < nu; t++) { o->U[t] += o->uprev[t]; o->U[t] += -1.0f * o->uhat_prev[t]; }
            for (int t = 0; t < nx; t++) { o->X[t] = o->xcur[t]; o->X[t] += o->e[t]; }
            gemv_n(nu, nv, 1.0f, o->L, nu, o->V, 1.0f, o->U);
            gemv_n(nx, nu, 1.0f, o->B, nx, o->U, 1.0f, o->X);
        } else {
            const int pc0 = o->cum[s - 1];
            const int branching = (ns - o->nps[s - 1]) > 0;
            #pragma omp parallel for schedule(static)
            for (int j = 0; j < ns; j++) {
                const int i = c0 + j;
                const int par = branching ? o->ancestor[i] - 1 : pc0 + j;
                real *u = o->U + (size_t)i * nu, *x = o->X + (size_t)i * nx;
                const real *up = o->U + (size_t)par * nu, *uhp = o->uhat + (size_t)par * nu;
                const real *xp = o->X + (size_t)par * nx, *ei = o->e + (size_t)i * nx;
                if (branching) {
                    gemv_n(nu, nv, 1.0f, o->L, nu, o->V + (size_t)i * nv, 1.0f, u);
                    for (int t = 0; t < nu; t++) { real lv = up[t] + -1.0f * uhp[t]; u[t] = lv + u[t]; }
                    for (int t = 0; t < nx; t++) x[t] = ei[t];
                    gemv_n(nx, nu, 1.0f, o->B, nx, u, 1.0f, x);
                    for (int t = 0; t < nx; t++) x[t] = xp[t] + x[t];
                } else {
                    for (int t = 0; t < nu; t++) { u[t] += up[t]; u[t] += -1.0f * uhp[t]; }
                    gemv_n(nu, nv, 1.0f, o->L, nu, o->V + (size_t)i * nv, 1.0f, u);
                    for (int t = 0; t < nx; t++) { x[t] = xp[t]; x[t] += ei[t]; }
                    gemv_n(nx, nu, 1.0f, o->B, nx, u, 1.0f, x);
                }
            }
        }
    }
    /* Hx = sysF x, sysG u (SmpcController.cu:744-747) */
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nodes; i++) {
        for (int t = 0; t < nx; t++) {
            o->pri_xi[(size_t)i * 2 * nx + t] = o->s_x[(size_t)i * nx + t] * o->X[(size_t)i * nx + t];
            o->pri_xi[(size_t)i * 2 * nx + nx + t] = o->s_xs[(size_t)i * nx + t] * o->X[(size_t)i * nx + t];
        }
        for (int t = 0; t < nu; t++) o->pri_psi[(size_t)i * nu + t] = o->s_u[(size_t)i * nu + t] * o->U[(size_t)i * nu + t];
    }
    free(tq); free(tr);
}

static real clampf(real v, real lo, real hi) { if (v < lo) return lo; else if (v > hi) return hi; return v; }

/* SmpcController::proximalFunG (SmpcController.cu:759-835), incl. the scratch-clobber quirk */
void orc_prox(orc *o) {
    const int nx = o->nx, nu = o->nu, nodes = o->nodes;
    const real inv_lambda = 1 / o->step;
    union { unsigned u; f32 f; } up; up.u = 0x7F7F7F7Fu;   /* cudaMemset(.., 127, ..): Engine.cu:454-455 */
    const real xs_upper = up.f;
    size_t nxi = (size_t)nodes * 2 * nx;
    real *diff = fdup(NULL, nxi);
    double s1 = 0.0, s2 = 0.0;
    #pragma omp parallel for schedule(static) reduction(+:s1,s2)
    for (int i = 0; i < nodes; i++) {
        for (int t = 0; t < nx; t++) {
            size_t k1 = (size_t)i * 2 * nx + t, k2 = k1 + nx;
            real t1 = o->pri_xi[k1] + inv_lambda * o->acc_xi[k1];
            real t2 = o->pri_xi[k2] + inv_lambda * o->acc_xi[k2];
            real z1 = clampf(t1, o->sxmin[(size_t)i * nx + t], o->sxmax[(size_t)i * nx + t]);
            real z2 = clampf(t2, o->sxs[(size_t)i * nx + t], xs_upper);
            o->dual_xi[k1] = z1; o->dual_xi[k2] = z2;
            diff[k1] = t1 + -1.0f * z1; diff[k2] = t2 + -1.0f * z2;
            s1 += (double)diff[k1] * diff[k1]; s2 += (double)diff[k2] * diff[k2];
        }
        for (int t = 0; t < nu; t++) {
            size_t k = (size_t)i * nu + t;
            real tu = o->pri_psi[k] + inv_lambda * o->acc_psi[k];
            o->dual_psi[k] = clampf(tu, o->sumin[k], o->sumax[k]);
        }
    }
    real d1 = (real)sqrt(s1), d2 = (real)sqrt(s2);
    o->dist_x = d1; o->dist_xs = d2;
    if (d1 > inv_lambda * o->pen_x) {
        real sc = 1 - inv_lambda * o->pen_x / d1;
        for (int i = 0; i < nodes; i++) for (int t = 0; t < nx; t++) { size_t k = (size_t)i * 2 * nx + t; o->dual_xi[k] = o->dual_xi[k] + sc * diff[k]; }
        /* :800-802 -- scratch reused for g(xBox): node 0 copied, part 1 clamped, then diff -= z everywhere */
        memcpy(diff, o->dual_xi, (size_t)2 * nx * sizeof(real));
        for (int i = 0; i < nodes; i++) for (int t = 0; t < nx; t++) { size_t k = (size_t)i * 2 * nx + t; diff[k] = clampf(diff[k], o->sxmin[(size_t)i * nx + t], o->sxmax[(size_t)i * nx + t]); }
        for (size_t k = 0; k < nxi; k++) diff[k] += -1.0f * o->dual_xi[k];
    }
    if (d2 > inv_lambda * o->pen_xs) {
        real sc = 1 - inv_lambda * o->pen_xs / d2;
        for (int i = 0; i < nodes; i++) for (int t = 0; t < nx; t++) { size_t k = (size_t)i * 2 * nx + nx + t; o->dual_xi[k] = o->dual_xi[k] + sc * diff[k]; }
    }
    free(diff);
}

/* computeFixedPointResidual (:839-850) */
void orc_residual(orc *o) {
    size_t nxi = (size_t)o->nodes * 2 * o->nx, nps = (size_t)o->nodes * o->nu;
    for (size_t t = 0; t < nxi; t++) o->res_xi[t] = o->pri_xi[t] + -1.0f * o->dual_xi[t];
    for (size_t t = 0; t < nps; t++) o->res_psi[t] = o->pri_psi[t] + -1.0f * o->dual_psi[t];
}

/* dualUpdate, APG branch (:854-864) */
void orc_dual_update(orc *o) {
    size_t nxi = (size_t)o->nodes * 2 * o->nx, nps = (size_t)o->nodes * o->nu;
    for (size_t t = 0; t < nxi; t++) o->upd_xi[t] = o->acc_xi[t] + o->step * o->res_xi[t];
    for (size_t t = 0; t < nps; t++) o->upd_psi[t] = o->acc_psi[t] + o->step * o->res_psi[t];
}

/* updatePrimalInfeasibity (:1480-1496): signed value at the first arg-max-abs, max over the two blocks */
real orc_primal_infeasibility(orc *o) {
    size_t nxi = (size_t)o->nodes * 2 * o->nx, nps = (size_t)o->nodes * o->nu;
    size_t ix = 0, ip = 0;
    for (size_t t = 1; t < nxi; t++) if (r_abs(o->res_xi[t]) > r_abs(o->res_xi[ix])) ix = t;
    for (size_t t = 1; t < nps; t++) if (r_abs(o->res_psi[t]) > r_abs(o->res_psi[ip])) ip = t;
    real a = o->res_xi[ix], b = o->res_psi[ip];
    return a > b ? a : b;
}

/* lambda table exactly as the host loop computes it (SmpcController.cu:1505-1520): real theta, double update */
void orc_lambda_table(int iters, real *lambda_out) {
    f32 theta0 = 1, theta1 = 1;
    for (int k = 0; k < iters; k++) {
        lambda_out[k] = (f32)(theta1 * (1 / theta0 - 1));
        theta0 = theta1;
        theta1 = 0.5 * (sqrt(pow(theta1, 4) + 4 * pow(theta1, 2)) - pow(theta1, 2));
    }
}

/* SmpcController::algorithmApg (:1500-1525) */
void orc_apg(orc *o, int iters, real *primal_infs /* nullable */) {
    real *lam = (real *)xcalloc(iters, sizeof(real));
    orc_lambda_table(iters, lam);
    orc_apg_init(o);
    for (int k = 0; k < iters; k++) {
        orc_extrapolate(o, lam[k]);
        orc_solve_step(o);
        orc_prox(o);
        orc_residual(o);
        orc_dual_update(o);
        real pi = orc_primal_infeasibility(o);
        if (primal_infs) primal_infs[k] = pi;
    }
    free(lam);
}

/* ---- named buffer access ---- */
static real *orc_buf(orc *o, const char *name, size_t *count) {
    size_t n = (size_t)o->nodes; const int nx = o->nx, nu = o->nu, nv = o->nv, nd = o->nd;
#define B_(nm, ptr, cnt) if (!strcmp(name, nm)) { *count = (cnt); return (ptr); }
    B_("L", o->L, (size_t)nu * nv) B_("Lhat", o->Lhat, (size_t)nu * nd) B_("Wv", o->Wv, (size_t)nu * nv)
    B_("Rbar", o->Rbar, (size_t)nv * nv) B_("G", o->Bbar, (size_t)nv * nx)
    B_("s_u", o->s_u, n * nu) B_("s_x", o->s_x, n * nx) B_("s_xs", o->s_xs, n * nx)
    B_("xmin", o->sxmin, n * nx) B_("xmax", o->sxmax, n * nx) B_("xs", o->sxs, n * nx)
    B_("umin", o->sumin, n * nu) B_("umax", o->sumax, n * nu)
    B_("Omega", o->Omega, (size_t)o->fbn * nv * nv) B_("Theta", o->Theta, (size_t)o->fbn * nv * nx)
    B_("Phi", o->Phi, n * 2 * nv * nx) B_("Psi", o->Psi, n * nv * nu) B_("D", o->D, n * 2 * nv * nx) B_("F", o->F, n * nv * nu)
    B_("e", o->e, n * nx) B_("uhat", o->uhat, n * nu) B_("alpha", o->alpha, n * nu) B_("beta", o->beta, n * nv)
    B_("uhat_prev", o->uhat_prev, (size_t)nu)
    B_("X", o->X, n * nx) B_("U", o->U, n * nu) B_("V", o->V, n * nv)
    B_("xi", o->xi, n * 2 * nx) B_("psi", o->psi, n * nu)
    B_("update_xi", o->upd_xi, n * 2 * nx) B_("update_psi", o->upd_psi, n * nu)
    B_("accel_xi", o->acc_xi, n * 2 * nx) B_("accel_psi", o->acc_psi, n * nu)
    B_("primal_xi", o->pri_xi, n * 2 * nx) B_("primal_psi", o->pri_psi, n * nu)
    B_("dual_xi", o->dual_xi, n * 2 * nx) B_("dual_psi", o->dual_psi, n * nu)
    B_("res_xi", o->res_xi, n * 2 * nx) B_("res_psi", o->res_psi, n * nu)
#undef B_
    *count = 0; return NULL;
}
long orc_count(orc *o, const char *name) { size_t c; return orc_buf(o, name, &c) ? (long)c : -1; }
int orc_get(orc *o, const char *name, real *out, long count) {
    size_t c; real *p = orc_buf(o, name, &c);
    if (!p || (size_t)count > c) return 1;
    memcpy(out, p, (size_t)count * sizeof(real)); return 0;
}
int orc_set(orc *o, const char *name, const real *in, long count) {
    size_t c; real *p = orc_buf(o, name, &c);
    if (!p || (size_t)count > c) return 1;
    memcpy(p, in, (size_t)count * sizeof(real)); return 0;
}
int orc_final_branch_node(orc *o) { return o->fb; }
real orc_distance(orc *o, int which) { return which ? o->dist_xs : o->dist_x; }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
