/* tests/capi_replay.c — plain-C replay of the MEX call sequence of the reference against include/qspush.h.
 *
 * What a maintainer's MEX gateway (matlab/qspush_mex.c + qspush_ocp.m) issues per control period when NMPC_controller.solve
 * (NMPC_controller.m:329-423) runs on top of the drop-in, batch = 1, every array from HOST memory:
 *
 *     x0(4) = mod(x0(4), b) - b (x0(4) < 0)                                  :332   (host arithmetic, like MATLAB)
 *     set('constr_x0', x0)                                                   :334
 *     for k = 0..Hp-1: set('cost_y_ref', y_ref(:, min(idx + k, T)), k)       :343-346   ONE CALL PER STAGE
 *     set('cost_y_ref_e', y_ref_{Hp-1}(1:4), Hp)                             :348       (stage argument = Hp, as the reference passes it)
 *     first call: xtraj = 0, utraj = [u_n_lb; 0], ptraj = 0                  :351-355
 *     v_bound clip of utraj(:, 1), Euler rollout with per-stage clip         :357-380   (plant.evalModelVariableShape -> qspush_eval_dynamics,
 *                                                                                        update_tangential_velocity_bounds -> qspush_eval_v_bound)
 *     set('init_x'), set('init_u'), set('init_pi')                           :382-384
 *     solve()                                                                :389
 *     get('u'), get('x'), get('pi'), shift                                   :392-399
 *     u = get('u', 0); get_cost                                              :403, 420
 *     get('status'), get('sqp_iter')                                         helper.m:253, 264
 *     plant: x <- x + dt * evalModelVariableShape(x, u)                      helper.m:294, 307
 *
 * qspush_opts is filled from C (qspush_opts_default + fields), so the struct layout the library was compiled with is
 * exercised by a second compiler; sizeof / offsets are written to the output for the ctypes mirror to be checked against.
 *
 * usage: capi_replay <input.bin> <output.bin>      (formats: tests/test_capi_replay.py)
 * build: gcc -O1 -std=c99 -Iinclude tests/capi_replay.c -Luclv_qs_pushing_matlab_b200 -lqspush -Wl,-rpath,... -lm -o tests/capi_replay
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qspush.h"

#define CHK(call) do { int rc__ = (call); if (rc__ != QSPUSH_OK) { fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc__, qspush_last_error()); return 2; } } while (0)

static double rd(FILE* f) { double v = 0; if (fread(&v, sizeof v, 1, f) != 1) { fprintf(stderr, "short input\n"); exit(3); } return v; }
static void rdv(FILE* f, double* p, size_t n) { if (fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "short input\n"); exit(3); } }

/* MATLAB mod (the builtin's algorithm, [MATLAB-RECALL]) in single precision: mod(double, single) -> single */
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

int main(int argc, char** argv) {
    if (argc == 2 && !strcmp(argv[1], "--layout")) {               /* struct layout as this compiler sees it (no GPU needed) */
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(qspush_opts), offsetof(qspush_opts, qp_max_iter), offsetof(qspush_opts, qp_tau),
               offsetof(qspush_opts, globalization), offsetof(qspush_opts, h_variant), offsetof(qspush_opts, qp_tol_comp),
               offsetof(qspush_opts, qp_stall), sizeof(qspush_ctrl), sizeof(qspush_loop_opts));
        return 0;
    }
    if (argc < 3) { fprintf(stderr, "usage: %s input.bin output.bin | --layout\n", argv[0]); return 1; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 1; }
    const int N = (int)rd(fi); const double dt = rd(fi); const int mode = (int)rd(fi); const int steps = (int)rd(fi); const int T = (int)rd(fi);
    const int nknots = (int)rd(fi), nctrl = (int)rd(fi);
    const double mu_sp = rd(fi), c_ellipse = rd(fi);
    double* knots = malloc(sizeof(double) * nknots); rdv(fi, knots, nknots);
    double* ctrl = malloc(sizeof(double) * 2 * nctrl); rdv(fi, ctrl, 2 * (size_t)nctrl);
    double W[36], We[16], x[4];
    rdv(fi, W, 36); rdv(fi, We, 16);
    double* yref = malloc(sizeof(double) * 6 * T); rdv(fi, yref, 6 * (size_t)T);     /* 6 x T column-major == [T][6] */
    rdv(fi, x, 4);
    fclose(fi);

    qspush_model* m = NULL;
    CHK(qspush_model_create(knots, nknots, ctrl, nctrl, 3, mu_sp, c_ellipse, 1, &m));
    double bb = 0, ce = 0, mu = 0; int nn = 0, nk = 0;
    CHK(qspush_model_info(m, &nn, &nk, &bb, &ce, &mu));
    qspush_opts o;
    qspush_opts_default(&o);
    o.mode = mode ? QSPUSH_MODE_SQP : QSPUSH_MODE_RTI;                     /* nlp_solver = "sqp" / "sqp_rti"   NMPC_controller.m:272 */
    o.max_sqp_iter = 30; o.tol_stat = o.tol_eq = o.tol_ineq = o.tol_comp = 1e-6;   /* :276 */
    o.globalization = 1;                                                   /* merit_backtracking              :272 */
    qspush_ctrl cc;
    qspush_ctrl_default(&cc);
    qspush_solver* s = NULL;
    const qspush_model* mm = m;
    CHK(qspush_solver_create(&mm, 1, N, dt, 1, 0, &o, &s));
    for (int k = 0; k < N; ++k) CHK(qspush_set(s, QSPUSH_W, k, 0, 0, W, QSPUSH_MEM_HOST));      /* update_cost_function :153-164 */
    CHK(qspush_set(s, QSPUSH_W, N, 0, 0, We, QSPUSH_MEM_HOST));

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) { perror(argv[2]); return 1; }
    const double layout[8] = {(double)sizeof(qspush_opts), (double)offsetof(qspush_opts, qp_max_iter), (double)offsetof(qspush_opts, qp_tau),
                              (double)offsetof(qspush_opts, globalization), (double)offsetof(qspush_opts, h_variant),
                              (double)offsetof(qspush_opts, qp_tol_comp), (double)offsetof(qspush_opts, qp_stall), (double)sizeof(qspush_ctrl)};
    fwrite(layout, sizeof(double), 8, fo);

    double* X = calloc((size_t)(N + 1) * 4, sizeof(double));      /* xtraj 4 x (N+1) */
    double* U = calloc((size_t)N * 2, sizeof(double));            /* utraj 2 x N */
    double* P = calloc((size_t)N * 4, sizeof(double));            /* ptraj 4 x N */
    int cold = 1;
    for (int i = 1; i <= steps; ++i) {
        double x0[4] = {x[0], x[1], x[2], x[3]};
        {   /* :332 (single arithmetic: b is a MATLAB single) */
            const double w = matlab_mod_single(x0[3], bb);
            x0[3] = (double)((float)w - (float)bb * (x0[3] < 0.0 ? 1.f : 0.f));
        }
        CHK(qspush_set(s, QSPUSH_X0, -1, 0, 1, x0, QSPUSH_MEM_HOST));                              /* :334 */
        for (int k = 0; k < N; ++k) {                                                              /* :343-346 */
            int col = i + k; if (col > T) col = T;
            CHK(qspush_set(s, QSPUSH_YREF, k, 0, 1, yref + (size_t)(col - 1) * 6, QSPUSH_MEM_HOST));
        }
        { int col = i + N - 1; if (col > T) col = T;
          CHK(qspush_set(s, QSPUSH_YREF_E, N, 0, 1, yref + (size_t)(col - 1) * 6, QSPUSH_MEM_HOST)); }   /* :348, stage = Hp like the reference */
        if (cold) {                                                                                /* :351-355 */
            memset(X, 0, sizeof(double) * (size_t)(N + 1) * 4); memset(P, 0, sizeof(double) * (size_t)N * 4);
            for (int k = 0; k < N; ++k) { U[2 * k] = cc.u_n_lb; U[2 * k + 1] = 0.0; }
            cold = 0;
        }
        double vb = 0, ta = 0;
        CHK(qspush_eval_v_bound(m, 0, QSPUSH_MEM_HOST, 1, &x0[3], &cc, 1, &vb, &ta));              /* :357 */
        for (int j = 0; j < 4; ++j) X[j] = x0[j];
        for (int k = 0; k < N; ++k) {                                                              /* :358-380 */
            if (fabs(U[2 * k + 1]) > vb) {
                const double old = U[2 * k + 1];
                U[2 * k + 1] = (double)((old > 0.0) - (old < 0.0)) * vb;
                U[2 * k] = U[2 * k + 1] * U[2 * k] / old;
            }
            double f[4];
            CHK(qspush_eval_dynamics(m, 0, QSPUSH_MEM_HOST, 1, X + 4 * k, U + 2 * k, f, NULL, NULL));   /* :369 */
            for (int j = 0; j < 4; ++j) X[4 * (k + 1) + j] = X[4 * k + j] + dt * f[j];
            if (k + 1 < N) CHK(qspush_eval_v_bound(m, 0, QSPUSH_MEM_HOST, 1, &X[4 * (k + 1) + 3], &cc, 1, &vb, &ta));   /* :371 */
        }
        CHK(qspush_set(s, QSPUSH_X, -1, 0, 1, X, QSPUSH_MEM_HOST));                                /* :382 */
        CHK(qspush_set(s, QSPUSH_U, -1, 0, 1, U, QSPUSH_MEM_HOST));                                /* :383 */
        CHK(qspush_set(s, QSPUSH_PI, -1, 0, 1, P, QSPUSH_MEM_HOST));                               /* :384 */
        CHK(qspush_solve(s));                                                                      /* :389 */
        CHK(qspush_get(s, QSPUSH_U, -1, 0, 1, U, QSPUSH_MEM_HOST));                                /* :392 */
        CHK(qspush_get(s, QSPUSH_X, -1, 0, 1, X, QSPUSH_MEM_HOST));                                /* :393 */
        CHK(qspush_get(s, QSPUSH_PI, -1, 0, 1, P, QSPUSH_MEM_HOST));                               /* :394 */
        double u0[2], cost = 0;
        CHK(qspush_get(s, QSPUSH_U, 0, 0, 1, u0, QSPUSH_MEM_HOST));                                /* :403 */
        CHK(qspush_get(s, QSPUSH_COST, -1, 0, 1, &cost, QSPUSH_MEM_HOST));                         /* :420 */
        int status = -1, sqp_iter = -1;
        CHK(qspush_get_int(s, QSPUSH_STATUS, 0, 1, &status, QSPUSH_MEM_HOST));                     /* helper.m:253 */
        CHK(qspush_get_int(s, QSPUSH_SQP_ITER, 0, 1, &sqp_iter, QSPUSH_MEM_HOST));                 /* helper.m:264 */
        /* :397-399 shift left, duplicate the last column */
        memmove(X, X + 4, sizeof(double) * (size_t)N * 4);
        memmove(U, U + 2, sizeof(double) * (size_t)(N - 1) * 2);
        memmove(P, P + 4, sizeof(double) * (size_t)(N - 1) * 4);
        const double rec[10] = {x[0], x[1], x[2], x[3], u0[0], u0[1], (double)status, (double)sqp_iter, cost, x0[3]};
        fwrite(rec, sizeof(double), 10, fo);
        double f[4];
        CHK(qspush_eval_dynamics(m, 0, QSPUSH_MEM_HOST, 1, x, u0, f, NULL, NULL));                 /* helper.m:294 */
        for (int j = 0; j < 4; ++j) x[j] += dt * f[j];                                             /* helper.m:307 */
    }
    fclose(fo);
    qspush_solver_free(s);
    qspush_model_free(m);
    free(knots); free(ctrl); free(yref); free(X); free(U); free(P);
    return 0;
}
