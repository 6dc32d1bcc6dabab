/* Minimal MEX API for matlab/qspush_mex.c in a container without MATLAB: the declarations the gateway uses (so that it is
 * type-checked against include/qspush.h) plus the few constructors a caller needs; tests/stubs/mex_runtime.c implements them
 * (column-major numeric arrays, char row vectors, logical scalars) so that the gateway can be EXECUTED by tests/mex_replay.c.
 * Test infrastructure only. */
#ifndef QSPUSH_TEST_MEX_STUB_H
#define QSPUSH_TEST_MEX_STUB_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxLOGICAL_CLASS = 3, mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT64_CLASS = 15 } mxClassID;
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
int mxGetString(const mxArray* a, char* buf, mwSize buflen);
void* mxGetData(const mxArray* a);
double* mxGetPr(const mxArray* a);
double mxGetScalar(const mxArray* a);
mwSize mxGetNumberOfDimensions(const mxArray* a);
const mwSize* mxGetDimensions(const mxArray* a);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity c);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateLogicalScalar(int v);
mxArray* mxCreateString(const char* s);
void mxDestroyArray(mxArray* a);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#endif
