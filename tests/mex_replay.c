/* tests/mex_replay.c — the MEX gateway EXECUTED outside MATLAB: matlab/qspush_mex.c is compiled against the working miniature
 * of the MEX array API in tests/stubs/mex_runtime.c, and this driver issues the mexFunction calls that matlab/qspush_ocp.m
 * issues when the unmodified NMPC_controller.solve (NMPC_controller.m:329-423) and helper.closed_loop_matlab
 * (helper.m:219-313) run on top of the drop-in: config 1 of BASELINE.json (main.m: one slider, x0 = 0, Hp = 10, 201 periods).
 *
 * The functions ocp_new / ocp_set / ocp_solve / ocp_get / ocp_get_cost / ocp_delete below are qspush_ocp.m transcribed method
 * by method (same argument lists, same field table F, same defaulting of `stage`); the control period is NMPC_controller.solve
 * line by line (cited).  The same period is then replayed through include/qspush.h directly (what tests/capi_replay.c does)
 * with the options the gateway sets, and the two runs must agree BIT FOR BIT on states, u0, status, sqp_iter and cost of every
 * period: the gateway adds nothing and loses nothing (layouts, batch inference from the array dimensions, handles, the
 * `stage = Hp` call of :348, scalar conversions).  What is NOT executed here is the MATLAB interpreter itself.
 *
 * usage: mex_replay <outline.ply> <mode 0|1> <steps> <output.bin>
 * build: gcc -O1 -std=c99 -Itests/stubs -Iinclude tests/mex_replay.c tests/stubs/mex_runtime.c matlab/qspush_mex.c -L... -lqspush -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mex.h"
#include "qspush.h"

#define CHK(call) do { int rc__ = (call); if (rc__ != QSPUSH_OK) { fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc__, qspush_last_error()); exit(2); } } while (0)

/* ---------------------------------------------------------------- temporaries of one mexFunction call */
static mxArray* g_tmp[16];
static int g_ntmp = 0;
static mxArray* T_(mxArray* a) { g_tmp[g_ntmp++] = a; return a; }
static mxArray* S_(const char* s) { return T_(mxCreateString(s)); }
static mxArray* D_(double v) { return T_(mxCreateDoubleScalar(v)); }
static void mex(int nlhs, mxArray** plhs, int nrhs, const mxArray* a0, const mxArray* a1, const mxArray* a2, const mxArray* a3,
                const mxArray* a4, const mxArray* a5, const mxArray* a6, const mxArray* a7) {
    const mxArray* prhs[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
    mxArray* none[1] = {NULL};
    mexFunction(nlhs, plhs ? plhs : none, nrhs, prhs);
    for (int i = 0; i < g_ntmp; ++i) mxDestroyArray(g_tmp[i]);
    g_ntmp = 0;
}

/* ---------------------------------------------------------------- qspush_ocp.m, method by method */
typedef struct { mxArray* s; mxArray* m; int N; } ocp_t;                       /* properties s; m; N          qspush_ocp.m:7-9 */
typedef struct { const char* name; int id; } fent;
static const fent F[] = {{"constr_x0", 0}, {"cost_y_ref", 1}, {"cost_y_ref_e", 2}, {"init_x", 3}, {"init_u", 4}, {"init_pi", 5},   /* :11-12 */
                         {"x", 3}, {"u", 4}, {"pi", 5}, {"cost_W", 16}, {"constr_lh", 17}, {"constr_uh", 18}, {"status", 32}, {"sqp_iter", 33}};
static int Fid(const char* f) {
    for (size_t i = 0; i < sizeof F / sizeof F[0]; ++i) if (!strcmp(F[i].name, f)) return F[i].id;
    fprintf(stderr, "Unrecognized field name \"%s\".\n", f); exit(4);          /* what self.F.(field) raises in MATLAB */
}
/* function self = qspush_ocp(plant, Hp, sample_time, nlp_solver)              :15-20 */
static void ocp_new(ocp_t* self, const char* pcl_path, int flip, double mu_sg, double mu_sp, double m, double tau_max, int Hp, double sample_time,
                    const char* nlp_solver) {
    mxArray* out[1] = {NULL};
    mex(1, out, 8, S_("model_from_ply"), S_(pcl_path), T_(mxCreateLogicalScalar(flip)), D_(3), D_(mu_sg), D_(mu_sp), D_(m), D_(tau_max));
    self->m = out[0];
    self->N = Hp;
    mex(1, out, 7, S_("solver_create"), self->m, D_(Hp), D_(sample_time), D_(1), D_(0), D_((double)!strcmp(nlp_solver, "sqp")), NULL);
    self->s = out[0];
}
/* function set(self, field, value, stage)                                      :21-27 ; nargs = nargin - 1 */
static void ocp_set(ocp_t* self, const char* field, const mxArray* value, int nargs, int stage) {
    if (nargs < 3) stage = -1;
    if (!strcmp(field, "constr_x0") || !strcmp(field, "cost_y_ref_e")) stage = -1;
    mex(0, NULL, 5, S_("set"), self->s, D_(Fid(field)), D_(stage), value, NULL, NULL, NULL);
}
/* function solve(self)                                                         :28 */
static void ocp_solve(ocp_t* self) { mex(0, NULL, 2, S_("solve"), self->s, NULL, NULL, NULL, NULL, NULL, NULL); }
/* function v = get(self, field, stage)                                         :29-41 ; nargs = nargin - 1 */
static mxArray* ocp_get(ocp_t* self, const char* field, int nargs, int stage) {
    mxArray* out[1] = {NULL};
    if (!strcmp(field, "status") || !strcmp(field, "sqp_iter")) mex(1, out, 3, S_("get_int"), self->s, D_(Fid(field)), NULL, NULL, NULL, NULL, NULL);
    else if (!strcmp(field, "time_tot")) mex(1, out, 3, S_("stat"), self->s, D_(0), NULL, NULL, NULL, NULL, NULL);
    else if (!strcmp(field, "time_lin")) mex(1, out, 3, S_("stat"), self->s, D_(1), NULL, NULL, NULL, NULL, NULL);
    else if (!strcmp(field, "time_qp_sol")) mex(1, out, 3, S_("stat"), self->s, D_(2), NULL, NULL, NULL, NULL, NULL);
    else {
        int d1, d2;
        if (!strcmp(field, "u")) { d1 = 2; d2 = self->N; } else if (!strcmp(field, "x")) { d1 = 4; d2 = self->N + 1; }
        else if (!strcmp(field, "pi")) { d1 = 4; d2 = self->N; } else { fprintf(stderr, "Unrecognized field name \"%s\".\n", field); exit(4); }
        if (nargs < 2) mex(1, out, 6, S_("get"), self->s, D_(Fid(field)), D_(-1), D_(d1), D_(d2), NULL, NULL);
        else mex(1, out, 6, S_("get"), self->s, D_(Fid(field)), D_(stage), D_(d1), D_(1), NULL, NULL);
    }
    return out[0];
}
/* function c = get_cost(self)                                                  :42 */
static mxArray* ocp_get_cost(ocp_t* self) { mxArray* out[1] = {NULL}; mex(1, out, 6, S_("get"), self->s, D_(7), D_(-1), D_(1), D_(1), NULL, NULL); return out[0]; }
/* function delete(self)                                                        :46 */
static void ocp_delete(ocp_t* self) {
    mex(0, NULL, 2, S_("solver_free"), self->s, NULL, NULL, NULL, NULL, NULL, NULL);
    mex(0, NULL, 2, S_("model_free"), self->m, NULL, NULL, NULL, NULL, NULL, NULL);
    mxDestroyArray(self->s); mxDestroyArray(self->m);
}

/* ---------------------------------------------------------------- the two back ends of one control period */
typedef struct {
    int use_mex;
    ocp_t ocp;                    /* through the gateway */
    qspush_solver* s;             /* through include/qspush.h directly */
    int N;
} backend;

static mxArray* mat(const double* v, int rows, int cols) {
    mxArray* a = mxCreateDoubleMatrix((mwSize)rows, (mwSize)cols, mxREAL);
    memcpy(mxGetPr(a), v, sizeof(double) * (size_t)rows * cols);
    return a;
}
static void be_set(backend* b, const char* field, const double* v, int rows, int cols, int nargs, int stage) {
    if (b->use_mex) { mxArray* a = mat(v, rows, cols); ocp_set(&b->ocp, field, a, nargs, stage); mxDestroyArray(a); return; }
    const int f = Fid(field);
    CHK(qspush_set(b->s, (qspush_field)f, nargs < 3 ? -1 : stage, 0, f >= QSPUSH_W ? 0 : 1, v, QSPUSH_MEM_HOST));
}
static void be_get(backend* b, const char* field, int nargs, int stage, double* out, int n) {
    if (b->use_mex) { mxArray* a = ocp_get(&b->ocp, field, nargs, stage); memcpy(out, mxGetPr(a), sizeof(double) * (size_t)n); mxDestroyArray(a); return; }
    CHK(qspush_get(b->s, (qspush_field)Fid(field), nargs < 2 ? -1 : stage, 0, 1, out, QSPUSH_MEM_HOST));
}
static double be_cost(backend* b) {
    double c = 0;
    if (b->use_mex) { mxArray* a = ocp_get_cost(&b->ocp); c = mxGetScalar(a); mxDestroyArray(a); return c; }
    CHK(qspush_get(b->s, QSPUSH_COST, -1, 0, 1, &c, QSPUSH_MEM_HOST));
    return c;
}
static int be_int(backend* b, const char* field) {
    int v = 0;
    if (b->use_mex) { mxArray* a = ocp_get(&b->ocp, field, 1, 0); v = (int)mxGetScalar(a); mxDestroyArray(a); return v; }
    CHK(qspush_get_int(b->s, (qspush_field)Fid(field), 0, 1, &v, QSPUSH_MEM_HOST));
    return v;
}
static void be_solve(backend* b) { if (b->use_mex) ocp_solve(&b->ocp); else CHK(qspush_solve(b->s)); }

/* MATLAB mod in single precision (mod(double, single) -> single), as tests/capi_replay.c */
static double matlab_mod_single(double s, double b) {
    const float x = (float)s, y = (float)b;
    float r, q;
    int req0;
    if (y == 0.f) return (double)x;
    if (x == 0.f) return 0.0;
    r = fmodf(x, y);
    req0 = (r == 0.f);
    if (!req0 && y > floorf(y)) { q = fabsf(x / y); req0 = !(fabsf(q - floorf(q + 0.5f)) > 1.1920929e-7f * q); }
    if (req0) r = 0.f; else if ((x < 0.f) != (y < 0.f)) r += y;
    return (double)r;
}

/* helper.closed_loop_matlab over NMPC_controller.solve; rec: steps x 10 = [x(4), u0(2), status, sqp_iter, cost, wrapped s] */
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

static void closed_loop(backend* b, const qspush_model* plant, double bb, int N, double dt, int steps, int T, const double* yref, double* rec) {
    qspush_ctrl cc;
    qspush_ctrl_default(&cc);
    double x[4] = {0, 0, 0, 0};                                                                    /* main.m:53-56,78 (index 1) */
    double* X = calloc((size_t)(N + 1) * 4, sizeof(double));
    double* U = calloc((size_t)N * 2, sizeof(double));
    double* P = calloc((size_t)N * 4, sizeof(double));
    const double W[36] = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1e-3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1e-3, 0, 0, 0, 0, 0, 0, 1e-3};   /* main.m:82-84 */
    const double We[16] = {2e5, 0, 0, 0, 0, 2e5, 0, 0, 0, 0, 20, 0, 0, 0, 0, 0};
    for (int k = 0; k < N; ++k) be_set(b, "cost_W", W, 6, 6, 3, k);                                /* update_cost_function :153-164 */
    be_set(b, "cost_W", We, 4, 4, 3, N);
    int cold = 1;
    for (int i = 1; i <= steps; ++i) {
        double x0[4] = {x[0], x[1], x[2], x[3]};
        { const double w = matlab_mod_single(x0[3], bb); x0[3] = (double)((float)w - (float)bb * (x0[3] < 0.0 ? 1.f : 0.f)); }   /* :332 */
        be_set(b, "constr_x0", x0, 4, 1, 2, 0);                                                    /* :334  set('constr_x0', x0) */
        for (int k = 0; k < N; ++k) {                                                              /* :343-346 */
            int col = i + k; if (col > T) col = T;
            be_set(b, "cost_y_ref", yref + (size_t)(col - 1) * 6, 6, 1, 3, k);
        }
        { int col = i + N - 1; if (col > T) col = T; be_set(b, "cost_y_ref_e", yref + (size_t)(col - 1) * 6, 4, 1, 3, N); }   /* :348, stage = Hp */
        if (cold) {                                                                                /* :351-355 */
            memset(X, 0, sizeof(double) * (size_t)(N + 1) * 4); memset(P, 0, sizeof(double) * (size_t)N * 4);
            for (int k = 0; k < N; ++k) { U[2 * k] = cc.u_n_lb; U[2 * k + 1] = 0.0; }
            cold = 0;
        }
        double vb = 0, ta = 0;
        CHK(qspush_eval_v_bound(plant, 0, QSPUSH_MEM_HOST, 1, &x0[3], &cc, 1, &vb, &ta));          /* :357 */
        for (int j = 0; j < 4; ++j) X[j] = x0[j];
        for (int k = 0; k < N; ++k) {                                                              /* :358-380 */
            if (fabs(U[2 * k + 1]) > vb) {
                const double old = U[2 * k + 1];
                U[2 * k + 1] = (double)((old > 0.0) - (old < 0.0)) * vb;
                U[2 * k] = U[2 * k + 1] * U[2 * k] / old;
            }
            double f[4];
            CHK(qspush_eval_dynamics(plant, 0, QSPUSH_MEM_HOST, 1, X + 4 * k, U + 2 * k, f, NULL, NULL));
            for (int j = 0; j < 4; ++j) X[4 * (k + 1) + j] = X[4 * k + j] + dt * f[j];
            if (k + 1 < N) CHK(qspush_eval_v_bound(plant, 0, QSPUSH_MEM_HOST, 1, &X[4 * (k + 1) + 3], &cc, 1, &vb, &ta));
        }
        be_set(b, "init_x", X, 4, N + 1, 2, 0);                                                    /* :382 */
        be_set(b, "init_u", U, 2, N, 2, 0);                                                        /* :383 */
        be_set(b, "init_pi", P, 4, N, 2, 0);                                                       /* :384 */
        be_solve(b);                                                                               /* :389 */
        be_get(b, "u", 1, 0, U, 2 * N);                                                            /* :392 */
        be_get(b, "x", 1, 0, X, 4 * (N + 1));                                                      /* :393 */
        be_get(b, "pi", 1, 0, P, 4 * N);                                                           /* :394 */
        double u0[2];
        be_get(b, "u", 2, 0, u0, 2);                                                               /* :403 */
        const double cost = be_cost(b);                                                            /* :420 */
        const int status = be_int(b, "status"), sqp_iter = be_int(b, "sqp_iter");                  /* helper.m:253, 264 */
        memmove(X, X + 4, sizeof(double) * (size_t)N * 4);                                         /* :397-399 */
        memmove(U, U + 2, sizeof(double) * (size_t)(N - 1) * 2);
        memmove(P, P + 4, sizeof(double) * (size_t)(N - 1) * 4);
        double* r = rec + (size_t)(i - 1) * 10;
        r[0] = x[0]; r[1] = x[1]; r[2] = x[2]; r[3] = x[3]; r[4] = u0[0]; r[5] = u0[1]; r[6] = status; r[7] = sqp_iter; r[8] = cost; r[9] = x0[3];
        double f[4];
        CHK(qspush_eval_dynamics(plant, 0, QSPUSH_MEM_HOST, 1, x, u0, f, NULL, NULL));             /* helper.m:294 */
        for (int j = 0; j < 4; ++j) x[j] += dt * f[j];                                             /* helper.m:307 */
    }
    free(X); free(U); free(P);
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s outline.ply mode steps output.bin\n", argv[0]); return 1; }
    const char* ply = argv[1];
    const int mode = atoi(argv[2]), steps = atoi(argv[3]);
    const int N = 10, T = 201; const double dt = 0.05;                                             /* main.m:40-41, 105 */
    const double mu_sg = 0.32, mu_sp = 0.19, mass = 0.2875, tau_max = 0.0251;                      /* object_selection.m (santal) */
    double* yref = calloc((size_t)6 * T, sizeof(double));                                          /* main.m:150-178 (straight-line stand-in) */
    for (int c = 0; c < T; ++c) { const double v = 0.01 * (c * dt); yref[6 * c] = v < 0.10 ? v : 0.10; }
    qspush_model* plant = NULL;
    CHK(qspush_model_create_from_ply(ply, 0, 3, mu_sg, mu_sp, mass, tau_max, &plant));
    double bb = 0, ce = 0, mu = 0; int nn = 0, nk = 0;
    CHK(qspush_model_info(plant, &nn, &nk, &bb, &ce, &mu));
    double* recA = calloc((size_t)steps * 10, sizeof(double));
    double* recB = calloc((size_t)steps * 10, sizeof(double));

    backend A; memset(&A, 0, sizeof A); A.use_mex = 1; A.N = N;
    ocp_new(&A.ocp, ply, 0, mu_sg, mu_sp, mass, tau_max, N, dt, mode ? "sqp" : "sqp_rti");
    const double tA0 = now_s();
    closed_loop(&A, plant, bb, N, dt, steps, T, yref, recA);
    const double tA = now_s() - tA0;
    /* single-stage field with an explicit stage through the raw gateway (what an older qspush_ocp.m sent): accepted */
    { mxArray* v = mat(yref, 4, 1); mex(0, NULL, 5, S_("set"), A.ocp.s, D_(2), D_(N), v, NULL, NULL, NULL); mxDestroyArray(v); }
    { mxArray* t = ocp_get(&A.ocp, "time_tot", 1, 0); if (!(mxGetScalar(t) > 0.0)) { fprintf(stderr, "time_tot not positive\n"); return 6; } mxDestroyArray(t); }
    ocp_delete(&A.ocp);

    backend B; memset(&B, 0, sizeof B); B.N = N;
    qspush_model* mB = NULL;
    CHK(qspush_model_create_from_ply(ply, 0, 3, mu_sg, mu_sp, mass, tau_max, &mB));
    qspush_opts o; qspush_opts_default(&o); o.mode = mode;                                         /* what the gateway's solver_create sets */
    const qspush_model* mm = mB;
    CHK(qspush_solver_create(&mm, 1, N, dt, 1, 0, &o, &B.s));
    const double tB0 = now_s();
    closed_loop(&B, plant, bb, N, dt, steps, T, yref, recB);
    const double tB = now_s() - tB0;
    qspush_solver_free(B.s); qspush_model_free(mB);

    const int same = memcmp(recA, recB, sizeof(double) * (size_t)steps * 10) == 0;
    FILE* fo = fopen(argv[4], "wb");
    if (!fo) { perror(argv[4]); return 1; }
    fwrite(recA, sizeof(double), (size_t)steps * 10, fo);
    fwrite(recB, sizeof(double), (size_t)steps * 10, fo);
    fclose(fo);
    printf("mex_replay: %d periods through mexFunction, %s the direct C-ABI run; final x = %.6f m, status(last) = %d\n", steps,
           same ? "bit-identical to" : "DIFFERENT from", recA[(size_t)(steps - 1) * 10], (int)recA[(size_t)(steps - 1) * 10 + 6]);
    printf("mex_replay: wall time per control period (host rollout through qspush_eval_*, %d field calls, solve, gets): %.3f ms through mexFunction, %.3f ms through the C-ABI\n",
           N + 12, 1e3 * tA / steps, 1e3 * tB / steps);
    qspush_model_free(plant);
    free(yref); free(recA); free(recB);
    return same ? 0 : 6;
}
