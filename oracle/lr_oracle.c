/*
 * oracle/lr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the numerical hot path of
 * linkedin/gdmix's random-effect / fixed-effect logistic-regression trainer.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object; the product path
 * (gdmix_b200/) never does.
 *
 * What is restated, with the reference location each function follows
 * (paths relative to /root/reference/gdmix-trainer/src/gdmix/):
 *
 *   re_loss_grad()   models/custom/binary_logistic_regression.py:84-110 (_loss),
 *                    :121-131 (_gradient), :73-82 / :112-119 (L2 terms),
 *                    :133-142 (intercept column FIRST)
 *   lbfgsb_*()       scipy.optimize.fmin_l_bfgs_b as called at
 *                    binary_logistic_regression.py:223-231 and
 *                    fixed_effect_lr_lbfgs_model.py:635-643.  scipy is a
 *                    third-party dependency that is NOT vendored in the
 *                    reference (setup.py pins scipy==1.5.4; this image has
 *                    1.18.1).  Its algorithm is L-BFGS-B 3.0 (Byrd, Lu, Nocedal,
 *                    Zhu; Morales & Nocedal 2011) with the MINPACK-2 line search
 *                    dcsrch/dcstep (More' & Thuente 1994).  With no bounds the
 *                    subspace step equals the classical L-BFGS two-loop
 *                    direction with H0 = I/theta, which is what is restated here
 *                    together with L-BFGS-B's exact driver logic (first step
 *                    1/||d||, skip rule, restart-on-line-search-failure, stop
 *                    tests and the scipy wrapper's maxiter/maxfun handling).
 *   re_variance()    binary_logistic_regression.py:144-189 (SIMPLE / FULL)
 *   re_score()       binary_logistic_regression.py:241-262, job_consumers.py:138-152
 *   threshold()      util/model_utils.py:4-12
 *   fe_loss_grad()   models/custom/fixed_effect_lr_lbfgs_model.py:309-392
 *                    (intercept LAST, not divided by n, l2 = 0.5*sum(x^2))
 *   java_string_hash / partition_id
 *                    gdmix-data/src/main/scala/com/linkedin/gdmix/utils/PartitionUtils.scala:31-37
 *
 * Parity pin: tests/golden/ holds vectors produced by running the reference's
 * own BinaryLogisticRegressionTrainer (imported from /root/reference) under
 * scipy 1.18.1 -- see oracle/gen_golden.py.  tests/test_oracle.py checks this
 * file against every one of them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

typedef struct {
    double l2;               /* lambda */
    int32_t regularize_bias; /* bool */
    int32_t has_intercept;   /* bool */
    int32_t m;               /* number of curvature pairs */
    int32_t max_iter;        /* scipy maxiter */
    int32_t max_ls;          /* scipy maxls (20) */
    int32_t max_fun;         /* scipy maxfun (15000) */
    double factr;            /* tolerance / eps */
    double pgtol;            /* 1e-5 (scipy default; the reference never sets it) */
} oracle_opts;

/* One entity's sample block in entity-local CSR form (what prepare_jobs builds,
 * job_consumers.py:243-247, before the ones column is prepended). */
typedef struct {
    int64_t n;             /* samples */
    int64_t d;             /* local features (without intercept) */
    const int64_t *rowptr; /* [n+1] */
    const int32_t *col;    /* [nnz] local feature index */
    const float *val;      /* [nnz] */
    const float *y;        /* [n] labels 0/1 */
    const float *w;        /* [n] */
    const float *off;      /* [n] */
} oracle_block;

/* numpy's pairwise summation (what ndarray.sum() does for a contiguous fp64
 * vector), so that cost.sum() is reproduced to the bit where possible. */
static double np_pairwise_sum(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int k = 0; k < 8; k++) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* ---- random-effect objective -------------------------------------------- */

/* theta has length d + has_intercept, intercept FIRST.  scratch: n doubles.
 * Returns f; writes g (same length as theta).  Follows the reference's COO
 * accumulation order: ones column first, then the row's features in order. */
static double re_loss_grad_impl(const oracle_block *b, const oracle_opts *o, const double *theta,
                                double *g, double *scratch)
{
    const int64_t n = b->n, d = b->d;
    const int hi = o->has_intercept ? 1 : 0;
    const int64_t p = d + hi;
    double *cost = scratch;      /* per-sample cost, then per-sample residual */
    for (int64_t j = 0; j < p; j++) g[j] = 0.0;

    /* z = X1.theta + offset ; cost_i ; r_i = w_i (sigmoid(z_i) - y_i) */
    double *resid = scratch + n;
    for (int64_t i = 0; i < n; i++) {
        double z = 0.0;
        if (hi) z += 1.0 * theta[0];
        for (int64_t k = b->rowptr[i]; k < b->rowptr[i + 1]; k++)
            z += (double)b->val[k] * theta[hi + b->col[k]];
        z = z + (double)b->off[i];
        double yi = (double)b->y[i], wi = (double)b->w[i];
        double ce = fmax(z, 0.0) - z * yi + log(1.0 + exp(-fabs(z)));
        cost[i] = wi * ce;
        double pr = 1.0 / (1.0 + exp(-z)); /* scipy.special.expit */
        resid[i] = wi * (pr - yi);
    }
    double reg = 0.0;
    {
        int64_t j0 = (hi && !o->regularize_bias) ? 1 : 0;
        for (int64_t j = j0; j < p; j++) reg += theta[j] * theta[j];
        reg *= o->l2 / 2.0;
    }
    double f = (1.0 / (double)n) * (np_pairwise_sum(cost, n) + reg);

    /* cost_grad = X1^T resid (COO order: intercept entries first, then row-major) */
    if (hi)
        for (int64_t i = 0; i < n; i++) g[0] += 1.0 * resid[i];
    for (int64_t i = 0; i < n; i++)
        for (int64_t k = b->rowptr[i]; k < b->rowptr[i + 1]; k++)
            g[hi + b->col[k]] += (double)b->val[k] * resid[i];
    for (int64_t j = 0; j < p; j++) {
        double gr = o->l2 * theta[j];
        if (hi && !o->regularize_bias && j == 0) gr = 0.0;
        g[j] = (1.0 / (double)n) * (g[j] + gr);
    }
    return f;
}

ORACLE_API double oracle_re_loss_grad(const oracle_block *b, const oracle_opts *o, const double *theta, double *g)
{
    double *scratch = (double *)malloc(sizeof(double) * 2 * (size_t)(b->n > 0 ? b->n : 1));
    double f = re_loss_grad_impl(b, o, theta, g, scratch);
    free(scratch);
    return f;
}

/* ---- MINPACK-2 dcsrch / dcstep (More' & Thuente) ------------------------- */

typedef struct {
    int brackt, stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
} dcsrch_state;

enum { LS_START = 0, LS_FG = 1, LS_CONV = 2, LS_WARN = 3, LS_ERROR = 4 };

static double dmax3(double a, double b, double c) { return fmax(fmax(a, b), c); }

static void dcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp,
                   double fp, double dp, int *brackt, double stpmin, double stpmax)
{
    double sgnd = dp * (*dx / fabs(*dx));
    double theta, s, gamma, p, q, r, stpc, stpq, stpf;

    if (fp > *fx) {
        /* case 1: higher function value -- minimum is bracketed */
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = dmax3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp < *stx) gamma = -gamma;
        p = (gamma - *dx) + theta;
        q = ((gamma - *dx) + gamma) + dp;
        r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.0) * (*stp - *stx);
        if (fabs(stpc - *stx) < fabs(stpq - *stx))
            stpf = stpc;
        else
            stpf = stpc + (stpq - stpc) / 2.0;
        *brackt = 1;
    } else if (sgnd < 0.0) {
        /* case 2: lower value, derivatives of opposite sign -- bracketed */
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = dmax3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + *dx;
        r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (fabs(stpc - *stp) > fabs(stpq - *stp))
            stpf = stpc;
        else
            stpf = stpq;
        *brackt = 1;
    } else if (fabs(dp) < fabs(*dx)) {
        /* case 3: lower value, same sign, derivative magnitude decreases */
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = dmax3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (*dx / s) * (dp / s)));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (*dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0)
            stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx)
            stpc = stpmax;
        else
            stpc = stpmin;
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (*brackt) {
            if (fabs(stpc - *stp) < fabs(stpq - *stp))
                stpf = stpc;
            else
                stpf = stpq;
            if (*stp > *stx)
                stpf = fmin(*stp + 0.66 * (*sty - *stp), stpf);
            else
                stpf = fmax(*stp + 0.66 * (*sty - *stp), stpf);
        } else {
            if (fabs(stpc - *stp) > fabs(stpq - *stp))
                stpf = stpc;
            else
                stpf = stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        /* case 4: lower value, same sign, derivative does not decrease */
        if (*brackt) {
            theta = 3.0 * (fp - *fy) / (*sty - *stp) + *dy + dp;
            s = dmax3(fabs(theta), fabs(*dy), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
            if (*stp > *sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + *dy;
            r = p / q;
            stpc = *stp + r * (*sty - *stp);
            stpf = stpc;
        } else if (*stp > *stx)
            stpf = stpmax;
        else
            stpf = stpmin;
    }

    /* update the interval which contains a minimizer */
    if (fp > *fx) {
        *sty = *stp; *fy = fp; *dy = dp;
    } else {
        if (sgnd < 0.0) { *sty = *stx; *fy = *fx; *dy = *dx; }
        *stx = *stp; *fx = fp; *dx = dp;
    }
    *stp = stpf;
}

/* One reverse-communication call. task in/out is one of LS_*. */
static int dcsrch(double *stp, double f, double g, double ftol, double gtol, double xtol, double stpmin,
                  double stpmax, int task, dcsrch_state *S)
{
    const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
    if (task == LS_START) {
        if (*stp < stpmin || *stp > stpmax || g >= 0.0 || stpmax < stpmin) return LS_ERROR;
        S->brackt = 0; S->stage = 1;
        S->finit = f; S->ginit = g; S->gtest = ftol * g;
        S->width = stpmax - stpmin; S->width1 = S->width / p5;
        S->stx = 0.0; S->fx = f; S->gx = g;
        S->sty = 0.0; S->fy = f; S->gy = g;
        S->stmin = 0.0; S->stmax = *stp + xtrapu * *stp;
        return LS_FG;
    }
    double ftest = S->finit + *stp * S->gtest;
    if (S->stage == 1 && f <= ftest && g >= 0.0) S->stage = 2;

    int out = LS_FG;
    if (S->brackt && (*stp <= S->stmin || *stp >= S->stmax)) out = LS_WARN;
    if (S->brackt && S->stmax - S->stmin <= xtol * S->stmax) out = LS_WARN;
    if (*stp == stpmax && f <= ftest && g <= S->gtest) out = LS_WARN;
    if (*stp == stpmin && (f > ftest || g >= S->gtest)) out = LS_WARN;
    if (f <= ftest && fabs(g) <= gtol * (-S->ginit)) out = LS_CONV;
    if (out != LS_FG) return out;

    if (S->stage == 1 && f <= S->fx && f > ftest) {
        double fm = f - *stp * S->gtest;
        double fxm = S->fx - S->stx * S->gtest;
        double fym = S->fy - S->sty * S->gtest;
        double gm = g - S->gtest;
        double gxm = S->gx - S->gtest;
        double gym = S->gy - S->gtest;
        dcstep(&S->stx, &fxm, &gxm, &S->sty, &fym, &gym, stp, fm, gm, &S->brackt, S->stmin, S->stmax);
        S->fx = fxm + S->stx * S->gtest;
        S->fy = fym + S->sty * S->gtest;
        S->gx = gxm + S->gtest;
        S->gy = gym + S->gtest;
    } else {
        dcstep(&S->stx, &S->fx, &S->gx, &S->sty, &S->fy, &S->gy, stp, f, g, &S->brackt, S->stmin, S->stmax);
    }
    if (S->brackt) {
        if (fabs(S->sty - S->stx) >= p66 * S->width1) *stp = S->stx + p5 * (S->sty - S->stx);
        S->width1 = S->width;
        S->width = fabs(S->sty - S->stx);
    }
    if (S->brackt) {
        S->stmin = fmin(S->stx, S->sty);
        S->stmax = fmax(S->stx, S->sty);
    } else {
        S->stmin = *stp + xtrapl * (*stp - S->stx);
        S->stmax = *stp + xtrapu * (*stp - S->stx);
    }
    *stp = fmax(*stp, stpmin);
    *stp = fmin(*stp, stpmax);
    if ((S->brackt && (*stp <= S->stmin || *stp >= S->stmax)) ||
        (S->brackt && S->stmax - S->stmin <= xtol * S->stmax))
        *stp = S->stx;
    return LS_FG;
}

/* ---- L-BFGS-B driver, unbounded case, as scipy.fmin_l_bfgs_b runs it ------ */

typedef double (*objective_fn)(void *ctx, const double *x, double *g);

typedef struct {
    int32_t nit, nfev, status; /* status = scipy warnflag: 0 converged, 1 maxiter/maxfun, 2 abnormal */
    int32_t task;              /* 1 pgtol, 2 factr, 3 maxiter, 4 maxfun, 5 abnormal line search, 6 ascent dir */
    double f;
} lbfgsb_result;

static double ddot(int64_t n, const double *a, const double *b)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

static void lbfgsb_minimize(int64_t n, double *x, objective_fn fun, void *ctx, const oracle_opts *o,
                            lbfgsb_result *res, double *g_out)
{
    const double epsmch = 2.220446049250313e-16;
    const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmx = 1e10;
    const int m = o->m;
    double *g = (double *)calloc((size_t)n, sizeof(double));
    double *d = (double *)calloc((size_t)n, sizeof(double));
    double *t = (double *)calloc((size_t)n, sizeof(double)); /* x at start of line search */
    double *r = (double *)calloc((size_t)n, sizeof(double)); /* g at start of line search */
    double *q = (double *)calloc((size_t)n, sizeof(double));
    double *S = (double *)calloc((size_t)n * (size_t)(m > 0 ? m : 1), sizeof(double));
    double *Y = (double *)calloc((size_t)n * (size_t)(m > 0 ? m : 1), sizeof(double));
    double *rho = (double *)calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    double *alpha = (double *)calloc((size_t)(m > 0 ? m : 1), sizeof(double));
    int col = 0, head = 0; /* ring: oldest at head, col pairs stored */
    double theta = 1.0;
    int iter = 0, nfev = 0;

    double f = fun(ctx, x, g);
    nfev = 1;
    res->status = 0; res->task = 0;

    double sbgnrm = 0.0;
    for (int64_t i = 0; i < n; i++) sbgnrm = fmax(sbgnrm, fabs(g[i]));
    if (sbgnrm <= o->pgtol) { res->task = 1; goto done; }

    for (;;) {
        /* search direction d = -H g  (two-loop, H0 = I/theta) */
        for (int64_t i = 0; i < n; i++) q[i] = g[i];
        for (int k = col - 1; k >= 0; k--) {
            int s = (head + k) % m;
            alpha[s] = rho[s] * ddot(n, S + (size_t)s * n, q);
            for (int64_t i = 0; i < n; i++) q[i] -= alpha[s] * Y[(size_t)s * n + i];
        }
        for (int64_t i = 0; i < n; i++) q[i] = q[i] / theta;
        for (int k = 0; k < col; k++) {
            int s = (head + k) % m;
            double beta = rho[s] * ddot(n, Y + (size_t)s * n, q);
            for (int64_t i = 0; i < n; i++) q[i] += S[(size_t)s * n + i] * (alpha[s] - beta);
        }
        for (int64_t i = 0; i < n; i++) d[i] = -q[i];

        /* line search (lnsrlb) */
        double dtd = ddot(n, d, d), dnorm = sqrt(dtd);
        double stp = (iter == 0) ? fmin(1.0 / dnorm, stpmx) : 1.0;
        memcpy(t, x, sizeof(double) * (size_t)n);
        memcpy(r, g, sizeof(double) * (size_t)n);
        double fold = f, gd, gdold = 0.0;
        int ifun = 0, iback = 0, info = 0, lstask = LS_START;
        dcsrch_state ls;
        for (;;) {
            gd = ddot(n, g, d);
            if (ifun == 0) {
                gdold = gd;
                if (gd >= 0.0) { info = -4; break; }
            }
            lstask = dcsrch(&stp, f, gd, ftol, gtol, xtol, 0.0, stpmx, lstask, &ls);
            if (lstask == LS_CONV || lstask == LS_WARN) break;
            if (lstask == LS_ERROR) { info = -4; break; }
            ifun++; iback = ifun - 1;
            if (iback >= o->max_ls) break;
            for (int64_t i = 0; i < n; i++) x[i] = stp * d[i] + t[i];
            f = fun(ctx, x, g);
            nfev++;
        }
        if (info != 0 || iback >= o->max_ls) {
            /* restore the previous iterate */
            memcpy(x, t, sizeof(double) * (size_t)n);
            memcpy(g, r, sizeof(double) * (size_t)n);
            f = fold;
            if (col == 0) {
                /* abnormal termination */
                res->status = 2; res->task = (info != 0) ? 6 : 5;
                iter++;
                goto done;
            }
            col = 0; head = 0; theta = 1.0; /* refresh memory, restart with steepest descent */
            continue;
        }
        iter++;

        /* scipy wrapper, on NEW_X (checked before the solver's own stop tests) */
        if (iter >= o->max_iter) { res->status = 1; res->task = 3; goto done; }
        if (nfev > o->max_fun) { res->status = 1; res->task = 4; goto done; }

        sbgnrm = 0.0;
        for (int64_t i = 0; i < n; i++) sbgnrm = fmax(sbgnrm, fabs(g[i]));
        if (sbgnrm <= o->pgtol) { res->task = 1; goto done; }
        {
            double ddum = dmax3(fabs(fold), fabs(f), 1.0);
            if ((fold - f) <= epsmch * o->factr * ddum) { res->task = 2; goto done; }
        }

        /* BFGS pair: s = stp*d, y = g - g_old */
        double rr = 0.0, dr, ddum;
        for (int64_t i = 0; i < n; i++) { r[i] = g[i] - r[i]; rr += r[i] * r[i]; }
        if (stp == 1.0) {
            dr = gd - gdold; ddum = -gdold;
        } else {
            dr = (gd - gdold) * stp; ddum = -gdold * stp;
            for (int64_t i = 0; i < n; i++) d[i] *= stp;
        }
        if (dr <= epsmch * ddum) continue; /* skip the update */
        if (m > 0) {
            int slot;
            if (col < m) { slot = (head + col) % m; col++; }
            else { slot = head; head = (head + 1) % m; }
            memcpy(S + (size_t)slot * n, d, sizeof(double) * (size_t)n);
            memcpy(Y + (size_t)slot * n, r, sizeof(double) * (size_t)n);
            rho[slot] = 1.0 / dr;
            theta = rr / dr;
        }
    }
done:
    res->nit = iter; res->nfev = nfev; res->f = f;
    if (g_out) memcpy(g_out, g, sizeof(double) * (size_t)n);
    free(g); free(d); free(t); free(r); free(q); free(S); free(Y); free(rho); free(alpha);
}

typedef struct { const oracle_block *b; const oracle_opts *o; double *scratch; } re_ctx;

static double re_objective(void *vctx, const double *x, double *g)
{
    re_ctx *c = (re_ctx *)vctx;
    return re_loss_grad_impl(c->b, c->o, x, g, c->scratch);
}

/* BinaryLogisticRegressionTrainer.fit for one entity.  theta is in/out
 * (theta0 on entry -- zeros for a cold start).  info = {nit, nfev, status, task}. */
ORACLE_API void oracle_re_fit(const oracle_block *b, const oracle_opts *o, double *theta, double *f_out,
                              int32_t *info, double *g_out)
{
    re_ctx c;
    c.b = b; c.o = o;
    c.scratch = (double *)malloc(sizeof(double) * 2 * (size_t)(b->n > 0 ? b->n : 1));
    lbfgsb_result res;
    lbfgsb_minimize(b->d + (o->has_intercept ? 1 : 0), theta, re_objective, &c, o, &res, g_out);
    free(c.scratch);
    *f_out = res.f;
    info[0] = res.nit; info[1] = res.nfev; info[2] = res.status; info[3] = res.task;
}

/* Batched form used by the CPU baseline: entities [e0, e1) of a packed batch.
 * Layout mirrors include/gdmix_b200.h's gdmix_re_batch (host pointers). */
ORACLE_API void oracle_re_fit_batch(int64_t e0, int64_t e1, const int64_t *ent_rowptr, const int64_t *rowptr,
                                    const int32_t *col, const float *val, const float *y, const float *w,
                                    const float *off, const int64_t *theta_ptr, const oracle_opts *o,
                                    const double *theta0, double *theta_out, double *f_out, int32_t *nit,
                                    int32_t *nfev, int32_t *status)
{
    const int hi = o->has_intercept ? 1 : 0;
    for (int64_t e = e0; e < e1; e++) {
        int64_t r0 = ent_rowptr[e], r1 = ent_rowptr[e + 1];
        int64_t n = r1 - r0;
        int64_t p = theta_ptr[e + 1] - theta_ptr[e];
        int64_t *lrp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
        int64_t base = rowptr[r0];
        for (int64_t i = 0; i <= n; i++) lrp[i] = rowptr[r0 + i] - base;
        oracle_block b;
        b.n = n; b.d = p - hi; b.rowptr = lrp; b.col = col + base; b.val = val + base;
        b.y = y + r0; b.w = w + r0; b.off = off + r0;
        double *th = theta_out + theta_ptr[e];
        for (int64_t j = 0; j < p; j++) th[j] = theta0 ? theta0[theta_ptr[e] + j] : 0.0;
        int32_t info[4];
        double f;
        oracle_re_fit(&b, o, th, &f, info, NULL);
        if (f_out) f_out[e] = f;
        if (nit) nit[e] = info[0];
        if (nfev) nfev[e] = info[1];
        if (status) status[e] = info[2];
        free(lrp);
    }
}

/* ---- variance (binary_logistic_regression.py:144-189) -------------------- */

/* Gauss-Jordan inverse with partial pivoting; returns 0 on success. */
static int invert_dense(int64_t p, double *A, double *Ainv)
{
    for (int64_t i = 0; i < p; i++)
        for (int64_t j = 0; j < p; j++) Ainv[i * p + j] = (i == j) ? 1.0 : 0.0;
    for (int64_t c = 0; c < p; c++) {
        int64_t piv = c;
        for (int64_t i = c + 1; i < p; i++)
            if (fabs(A[i * p + c]) > fabs(A[piv * p + c])) piv = i;
        if (A[piv * p + c] == 0.0) return -1;
        if (piv != c)
            for (int64_t j = 0; j < p; j++) {
                double tmp = A[c * p + j]; A[c * p + j] = A[piv * p + j]; A[piv * p + j] = tmp;
                tmp = Ainv[c * p + j]; Ainv[c * p + j] = Ainv[piv * p + j]; Ainv[piv * p + j] = tmp;
            }
        double inv = 1.0 / A[c * p + c];
        for (int64_t j = 0; j < p; j++) { A[c * p + j] *= inv; Ainv[c * p + j] *= inv; }
        for (int64_t i = 0; i < p; i++) {
            if (i == c) continue;
            double fct = A[i * p + c];
            if (fct == 0.0) continue;
            for (int64_t j = 0; j < p; j++) {
                A[i * p + j] -= fct * A[c * p + j];
                Ainv[i * p + j] -= fct * Ainv[c * p + j];
            }
        }
    }
    return 0;
}

/* mode 1 = SIMPLE, 2 = FULL.  var has length d + has_intercept. */
ORACLE_API int oracle_re_variance(const oracle_block *b, const oracle_opts *o, const double *theta, int mode,
                                  double *var)
{
    const int64_t n = b->n, d = b->d;
    const int hi = o->has_intercept ? 1 : 0;
    const int64_t p = d + hi;
    const double epsilon = 1.0e-12;
    double *dw = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        double z = hi ? theta[0] : 0.0;
        for (int64_t k = b->rowptr[i]; k < b->rowptr[i + 1]; k++)
            z += (double)b->val[k] * theta[hi + b->col[k]];
        z += (double)b->off[i];
        double rho = 1.0 / (1.0 + exp(-z));
        dw[i] = rho * (1.0 - rho) * (double)b->w[i];
    }
    int rc = 0;
    if (mode == 1) {
        for (int64_t j = 0; j < p; j++) var[j] = 0.0;
        for (int64_t i = 0; i < n; i++) {
            if (hi) var[0] += dw[i];
            for (int64_t k = b->rowptr[i]; k < b->rowptr[i + 1]; k++) {
                double v = (double)b->val[k];
                var[hi + b->col[k]] += v * (v * dw[i]);
            }
        }
        for (int64_t j = 0; j < p; j++) {
            double h = var[j] + o->l2;
            if (hi && !o->regularize_bias && j == 0) h -= o->l2;
            var[j] = 1.0 / (h + epsilon);
        }
    } else {
        double *H = (double *)calloc((size_t)(p * p), sizeof(double));
        double *Hi = (double *)calloc((size_t)(p * p), sizeof(double));
        double *row = (double *)calloc((size_t)p, sizeof(double));
        for (int64_t i = 0; i < n; i++) {
            for (int64_t j = 0; j < p; j++) row[j] = 0.0;
            if (hi) row[0] = 1.0;
            for (int64_t k = b->rowptr[i]; k < b->rowptr[i + 1]; k++) row[hi + b->col[k]] += (double)b->val[k];
            for (int64_t a = 0; a < p; a++) {
                if (row[a] == 0.0) continue;
                for (int64_t c = 0; c < p; c++) H[a * p + c] += row[a] * (row[c] * dw[i]);
            }
        }
        for (int64_t j = 0; j < p; j++) H[j * p + j] += o->l2 + epsilon;
        if (hi && !o->regularize_bias) H[0] -= o->l2;
        rc = invert_dense(p, H, Hi);
        for (int64_t j = 0; j < p; j++) var[j] = Hi[j * p + j];
        free(H); free(Hi); free(row);
    }
    free(dw);
    return rc;
}

/* ---- scoring (binary_logistic_regression.py:241-262, job_consumers.py:138-152) */

/* theta == NULL means "entity has no model": logits = offsets. */
ORACLE_API void oracle_re_score(const oracle_block *b, const oracle_opts *o, const double *theta, double *logit,
                                double *logit_per_coordinate)
{
    const int hi = o->has_intercept ? 1 : 0;
    for (int64_t i = 0; i < b->n; i++) {
        double offs = (double)b->off[i];
        double z;
        if (!theta) {
            z = offs;
        } else {
            z = hi ? theta[0] : 0.0;
            for (int64_t k = b->rowptr[i]; k < b->rowptr[i + 1]; k++)
                z += (double)b->val[k] * theta[hi + b->col[k]];
            z = z + offs;
        }
        logit[i] = z;
        logit_per_coordinate[i] = z - offs;
    }
}

/* util/model_utils.py:4-12 */
ORACLE_API void oracle_threshold(double *coef, int64_t n, double threshold)
{
    for (int64_t i = 0; i < n; i++)
        if (fabs(coef[i]) <= threshold) coef[i] = 0.0;
}

/* ---- fixed-effect objective (fixed_effect_lr_lbfgs_model.py:309-392) ----- */

typedef struct {
    int64_t n, D;          /* rows, features (x has D + has_intercept entries, intercept LAST) */
    const int64_t *rowptr;
    const int32_t *col;
    const float *val;
    const float *y, *w, *off;
    int32_t linear_regression; /* 0: logistic, 1: squared error */
    int32_t num_workers;       /* the reference adds l2-term / num_workers on each worker */
} oracle_fe_rows;

ORACLE_API double oracle_fe_loss_grad(const oracle_fe_rows *R, const oracle_opts *o, const double *x, double *g)
{
    const int hi = o->has_intercept ? 1 : 0;
    const int64_t D = R->D, p = D + hi;
    for (int64_t j = 0; j < p; j++) g[j] = 0.0;
    double value = 0.0;
    for (int64_t i = 0; i < R->n; i++) {
        double z = 0.0;
        for (int64_t k = R->rowptr[i]; k < R->rowptr[i + 1]; k++) z += (double)R->val[k] * x[R->col[k]];
        z += (double)R->off[i];
        if (hi) z += x[D];
        double yi = (double)R->y[i], wi = (double)R->w[i], dz;
        if (R->linear_regression) {
            double e = yi - z;
            value += wi * e * e;
            dz = -2.0 * wi * e;
        } else {
            value += wi * (fmax(z, 0.0) - z * yi + log1p(exp(-fabs(z))));
            dz = wi * (1.0 / (1.0 + exp(-z)) - yi);
        }
        for (int64_t k = R->rowptr[i]; k < R->rowptr[i + 1]; k++) g[R->col[k]] += (double)R->val[k] * dz;
        if (hi) g[D] += dz;
    }
    int64_t preg = (hi && !o->regularize_bias) ? D : p;
    double sq = 0.0;
    for (int64_t j = 0; j < preg; j++) sq += x[j] * x[j];
    double nw = (double)(R->num_workers > 0 ? R->num_workers : 1);
    value += o->l2 * 0.5 * sq / nw;
    for (int64_t j = 0; j < preg; j++) g[j] += o->l2 * x[j] / nw;
    return value;
}

typedef struct { const oracle_fe_rows *R; const oracle_opts *o; } fe_ctx;
static double fe_objective(void *vctx, const double *x, double *g)
{
    fe_ctx *c = (fe_ctx *)vctx;
    return oracle_fe_loss_grad(c->R, c->o, x, g);
}

/* Single-worker fixed-effect solve (x in/out, intercept last). */
ORACLE_API void oracle_fe_fit(const oracle_fe_rows *R, const oracle_opts *o, double *x, double *f_out,
                              int32_t *info)
{
    fe_ctx c; c.R = R; c.o = o;
    lbfgsb_result res;
    lbfgsb_minimize(R->D + (o->has_intercept ? 1 : 0), x, fe_objective, &c, o, &res, NULL);
    *f_out = res.f;
    info[0] = res.nit; info[1] = res.nfev; info[2] = res.status; info[3] = res.task;
}

/* ---- entity -> partition map (PartitionUtils.scala:31-37) ---------------- */

/* java.lang.String.hashCode over UTF-16 code units; the caller passes the code
 * units (Python side encodes utf-16-le), so non-BMP ids hash like the JVM. */
ORACLE_API int32_t oracle_java_string_hash(const uint16_t *units, int64_t n)
{
    uint32_t h = 0;
    for (int64_t i = 0; i < n; i++) h = 31u * h + (uint32_t)units[i];
    return (int32_t)h;
}

/* abs(hash) % numPartitions with JVM semantics: Math.abs(Int.MinValue) stays
 * negative and % keeps the dividend's sign. */
ORACLE_API int32_t oracle_partition_id(const uint16_t *units, int64_t n, int32_t num_partitions)
{
    int32_t h = oracle_java_string_hash(units, n);
    int32_t a = (h == INT32_MIN) ? h : (h < 0 ? -h : h);
    return a % num_partitions; /* C99 % truncates toward zero like the JVM */
}
