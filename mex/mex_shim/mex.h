/* mex.h -- PROTOTYPE SHIM for compile-checking the gateways where MATLAB/Octave is absent.
 * Declares only the subset of the MEX C API (R2018a interleaved-complex flavour) the gateways use.
 * The real MEX file is built with `mex -R2018a` (MATLAB) or `mkoctfile --mex` (Octave), which supply the genuine header
 * and library.  For the tests, mex_stub_runtime.cpp implements these entry points well enough to LINK the
 * gateways against librsp.so and RUN them (tests/test_mex_gateway.py). */
#ifndef RSP_MEX_SHIM_H_
#define RSP_MEX_SHIM_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef struct { double real, imag; } mxComplexDouble;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
int mxIsStruct(const mxArray*);
int mxIsDouble(const mxArray*);
int mxIsComplex(const mxArray*);
int mxIsEmpty(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
double mxGetScalar(const mxArray*);
double* mxGetDoubles(const mxArray*);
mxComplexDouble* mxGetComplexDoubles(const mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
mxArray* mxCreateDoubleScalar(double);
void mxSetField(mxArray*, mwIndex, const char*, mxArray*);
void mxDestroyArray(mxArray*);
void* mxMalloc(size_t);
void mxFree(void*);
int mexCallMATLAB(int, mxArray**, int, mxArray**, const char*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexPrintf(const char*, ...);
void mexLock(void);
int mexAtExit(void (*)(void));
#ifdef __cplusplus
}
#endif
#endif
