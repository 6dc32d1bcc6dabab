/* tests/stubs/mex_runtime.c — a working miniature of the MEX array API declared in tests/stubs/mex.h, enough to EXECUTE
 * matlab/qspush_mex.c outside MATLAB (tests/mex_replay.c).  Semantics follow the documented behaviour of the MATLAB C Matrix
 * API: arrays are column-major, a matrix has 2 dimensions, mxGetScalar converts the first element of any numeric / logical
 * array to double, mxGetString returns 0 on success, mexErrMsgIdAndTxt does not return.  Test infrastructure only. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

struct mxArray_tag {
    mxClassID cls;
    mwSize ndim;
    mwSize dims[4];
    size_t elsize;
    void* data;
};

static size_t class_size(mxClassID c) {
    switch (c) {
        case mxLOGICAL_CLASS: return 1;
        case mxCHAR_CLASS: return 1;          /* (MATLAB stores UTF-16; the gateway only reads strings through mxGetString) */
        case mxDOUBLE_CLASS: return 8;
        case mxINT32_CLASS: return 4;
        case mxUINT64_CLASS: return 8;
    }
    return 0;
}

mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity c) {
    if (c != mxREAL || ndim < 1 || ndim > 4 || !class_size(cls)) { fprintf(stderr, "mex_runtime: unsupported array\n"); exit(5); }
    mxArray* a = calloc(1, sizeof *a);
    size_t n = 1;
    a->cls = cls; a->ndim = ndim < 2 ? 2 : ndim; a->elsize = class_size(cls);
    for (mwSize i = 0; i < 4; ++i) a->dims[i] = 1;
    for (mwSize i = 0; i < ndim; ++i) { a->dims[i] = dims[i]; n *= dims[i]; }
    while (a->ndim > 2 && a->dims[a->ndim - 1] == 1) a->ndim--;      /* MATLAB drops trailing singleton dimensions */
    a->data = calloc(n ? n : 1, a->elsize);
    return a;
}
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c) { const mwSize d[2] = {m, n}; return mxCreateNumericArray(2, d, cls, c); }
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c); }
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); *(double*)a->data = v; return a; }
mxArray* mxCreateLogicalScalar(int v) { mxArray* a = mxCreateNumericMatrix(1, 1, mxLOGICAL_CLASS, mxREAL); *(unsigned char*)a->data = v ? 1 : 0; return a; }
mxArray* mxCreateString(const char* s) {
    const mwSize d[2] = {1, strlen(s)};
    mxArray* a = mxCreateNumericArray(2, d, mxCHAR_CLASS, mxREAL);
    memcpy(a->data, s, d[1]);
    return a;
}
void mxDestroyArray(mxArray* a) { if (a) { free(a->data); free(a); } }

void* mxGetData(const mxArray* a) { return a->data; }
double* mxGetPr(const mxArray* a) {
    if (a->cls != mxDOUBLE_CLASS) { fprintf(stderr, "mex_runtime: mxGetPr on a non-double array\n"); exit(5); }
    return (double*)a->data;
}
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->ndim; }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims; }
double mxGetScalar(const mxArray* a) {
    switch (a->cls) {
        case mxLOGICAL_CLASS: return (double)*(unsigned char*)a->data;
        case mxCHAR_CLASS: return (double)*(unsigned char*)a->data;
        case mxDOUBLE_CLASS: return *(double*)a->data;
        case mxINT32_CLASS: return (double)*(int*)a->data;
        case mxUINT64_CLASS: return (double)*(unsigned long long*)a->data;
    }
    return 0.0;
}
int mxGetString(const mxArray* a, char* buf, mwSize buflen) {
    if (a->cls != mxCHAR_CLASS) return 1;
    const mwSize n = a->dims[0] * a->dims[1];
    if (n + 1 > buflen) return 1;
    memcpy(buf, a->data, n);
    buf[n] = 0;
    return 0;
}
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "MEX error %s: ", id);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(4);                                                     /* like MATLAB: control does not return to the gateway */
}
