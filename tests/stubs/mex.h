/* Minimal declarations of the MATLAB MEX API used by matlab/qspush_mex.c, so that the gateway can be compiled
 * (syntax and type checked against include/qspush.h) in a container without MATLAB.  Test infrastructure only. */
#ifndef QSPUSH_TEST_MEX_STUB_H
#define QSPUSH_TEST_MEX_STUB_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT64_CLASS = 15 } mxClassID;
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
int mxGetString(const mxArray* a, char* buf, mwSize buflen);
void* mxGetData(const mxArray* a);
double* mxGetPr(const mxArray* a);
double mxGetScalar(const mxArray* a);
mwSize mxGetNumberOfDimensions(const mxArray* a);
const mwSize* mxGetDimensions(const mxArray* a);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#endif
