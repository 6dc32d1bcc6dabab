/* matlab/qspush_mex.c — MEX gateway over include/qspush.h (one gateway, string-dispatched).
 *
 * Build (MATLAB R2022b, on a box with the CUDA runtime):
 *     mex -I../include qspush_mex.c -L../uclv_qs_pushing_matlab_b200 -lqspush
 * Not compiled in this repository's CI: the build container has neither MATLAB nor Octave
 * (SURVEY.md section 8c).  It is the binding a maintainer of the reference adds; see INTEGRATION.md.
 *
 *   h  = qspush_mex('model_from_ply', ply_path, flip, degree, mu_sg, mu_sp, m, tau_max)
 *   s  = qspush_mex('solver_create', h_model, N, dt, batch, device, mode [, h_variant])   % mode 0 = sqp_rti, 1 = sqp;
 *                                                    % h_variant 1 = h = [u_n; u_t -+ v_bound(s)] (NMPC_controller.m:238)
 *        qspush_mex('set', s, field_id, stage, data)      % data: dim x stages x batch (column-major == C-ABI layout)
 *   v  = qspush_mex('get', s, field_id, stage, nrows, ncols)
 *   v  = qspush_mex('get_int', s, field_id)
 *        qspush_mex('prepare', s) ; qspush_mex('solve', s) ; qspush_mex('shift', s)
 *   t  = qspush_mex('stat', s, which)
 *   [x, xlog, ulog] = qspush_mex('closed_loop', s, traj(6 x T), x0(4 x batch), steps)   % device-resident helper.closed_loop_matlab
 *        qspush_mex('solver_free', s) ; qspush_mex('model_free', h)
 */
#include <string.h>
#include "mex.h"
#include "qspush.h"

static void chk(int rc) { if (rc != QSPUSH_OK) mexErrMsgIdAndTxt("qspush:error", "%s", qspush_last_error()); }
static void* hnd(const mxArray* a) { return (void*)(*(unsigned long long*)mxGetData(a)); }
static mxArray* mkh(void* p) { mxArray* a = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL); *(unsigned long long*)mxGetData(a) = (unsigned long long)p; return a; }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[64];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof cmd)) mexErrMsgIdAndTxt("qspush:usage", "first argument must be a command string");
    if (!strcmp(cmd, "model_from_ply")) {
        char path[4096]; mxGetString(prhs[1], path, sizeof path);
        qspush_model* m = NULL;
        chk(qspush_model_create_from_ply(path, (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3]), mxGetScalar(prhs[4]),
                                         mxGetScalar(prhs[5]), mxGetScalar(prhs[6]), mxGetScalar(prhs[7]), &m));
        plhs[0] = mkh(m);
    } else if (!strcmp(cmd, "solver_create")) {
        const qspush_model* m = (const qspush_model*)hnd(prhs[1]);
        qspush_opts o; qspush_opts_default(&o); o.mode = (int)mxGetScalar(prhs[6]);
        if (nrhs > 7) o.h_variant = (int)mxGetScalar(prhs[7]);
        qspush_solver* s = NULL;
        chk(qspush_solver_create(&m, 1, (int)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]), (int)mxGetScalar(prhs[5]), &o, &s));
        plhs[0] = mkh(s);
    } else if (!strcmp(cmd, "set")) {
        qspush_solver* s = (qspush_solver*)hnd(prhs[1]);
        int f = (int)mxGetScalar(prhs[2]), stage = (int)mxGetScalar(prhs[3]);
        int batch = 1; { const mwSize* d = mxGetDimensions(prhs[4]); if (mxGetNumberOfDimensions(prhs[4]) == 3) batch = (int)d[2]; }
        chk(qspush_set(s, (qspush_field)f, stage, 0, f >= QSPUSH_W ? 0 : batch, mxGetPr(prhs[4]), QSPUSH_MEM_HOST));
    } else if (!strcmp(cmd, "get")) {
        qspush_solver* s = (qspush_solver*)hnd(prhs[1]);
        int f = (int)mxGetScalar(prhs[2]), stage = (int)mxGetScalar(prhs[3]);
        plhs[0] = mxCreateDoubleMatrix((mwSize)mxGetScalar(prhs[4]), (mwSize)mxGetScalar(prhs[5]), mxREAL);
        chk(qspush_get(s, (qspush_field)f, stage, 0, 1, mxGetPr(plhs[0]), QSPUSH_MEM_HOST));
    } else if (!strcmp(cmd, "get_int")) {
        int v = 0; chk(qspush_get_int((qspush_solver*)hnd(prhs[1]), (qspush_field)(int)mxGetScalar(prhs[2]), 0, 1, &v, QSPUSH_MEM_HOST));
        plhs[0] = mxCreateDoubleScalar((double)v);
    } else if (!strcmp(cmd, "prepare")) chk(qspush_prepare((qspush_solver*)hnd(prhs[1])));
    else if (!strcmp(cmd, "solve")) chk(qspush_solve((qspush_solver*)hnd(prhs[1])));
    else if (!strcmp(cmd, "shift")) chk(qspush_shift((qspush_solver*)hnd(prhs[1])));
    else if (!strcmp(cmd, "set_reference")) {
        /* y_ref is 6 x T column-major == [T][6] (controller.y_ref, NMPC_controller.m:425-431); kept on the device */
        chk(qspush_set_reference_trajectory((qspush_solver*)hnd(prhs[1]), mxGetPr(prhs[2]), (int)mxGetDimensions(prhs[2])[1], NULL, QSPUSH_MEM_HOST));
    } else if (!strcmp(cmd, "reference_window")) chk(qspush_set_reference_window((qspush_solver*)hnd(prhs[1]), (int)mxGetScalar(prhs[2])));
    else if (!strcmp(cmd, "stat")) { double v = 0; chk(qspush_get_stat((qspush_solver*)hnd(prhs[1]), (qspush_stat)(int)mxGetScalar(prhs[2]), &v)); plhs[0] = mxCreateDoubleScalar(v); }
    else if (!strcmp(cmd, "closed_loop")) {
        /* traj is 6 x T column-major == [T][6]; x0 is 4 x batch == [batch][4]: the C-ABI layouts */
        qspush_solver* s = (qspush_solver*)hnd(prhs[1]);
        const mwSize T = mxGetDimensions(prhs[2])[1], batch = mxGetDimensions(prhs[3])[1];
        const int steps = (int)mxGetScalar(prhs[4]);
        qspush_loop_opts lo; memset(&lo, 0, sizeof lo); lo.idx0 = 1;
        plhs[0] = mxCreateDoubleMatrix(4, batch, mxREAL);
        memcpy(mxGetPr(plhs[0]), mxGetPr(prhs[3]), sizeof(double) * 4 * batch);
        mxArray* xl = mxCreateDoubleMatrix(4 * batch, (mwSize)steps, mxREAL);
        mxArray* ul = mxCreateDoubleMatrix(2 * batch, (mwSize)steps, mxREAL);
        chk(qspush_closed_loop(s, mxGetPr(prhs[2]), (int)T, NULL, mxGetPr(plhs[0]), steps, &lo, mxGetPr(xl), mxGetPr(ul), NULL, QSPUSH_MEM_HOST));
        if (nlhs > 1) plhs[1] = xl;
        if (nlhs > 2) plhs[2] = ul;
    }
    else if (!strcmp(cmd, "solver_free")) qspush_solver_free((qspush_solver*)hnd(prhs[1]));
    else if (!strcmp(cmd, "model_free")) qspush_model_free((qspush_model*)hnd(prhs[1]));
    else mexErrMsgIdAndTxt("qspush:usage", "unknown command %s", cmd);
}
