/* Minimal stand-in for MATLAB's mex.h / matrix.h, ONLY for syntax-checking bellman_mex.cpp in an
 * image that has no MATLAB (g++ -fsyntax-only -Istub).  A real build uses MATLAB's own headers:
 *     mex -R2018a bellman_mex.cpp -I../../include -L.. -lbellman
 * Declarations follow the documented C Matrix / MEX API. */
#ifndef STUB_MEX_H
#define STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS,
               mxDOUBLE_CLASS, mxSINGLE_CLASS, mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS,
               mxINT32_CLASS, mxUINT32_CLASS, mxINT64_CLASS, mxUINT64_CLASS } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexLock(void);
void mexUnlock(void);
int mexAtExit(void (*fn)(void));
int mexPrintf(const char *fmt, ...);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray *mxCreateString(const char *s);
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **names);
void mxSetField(mxArray *s, mwIndex i, const char *name, mxArray *v);
mxArray *mxGetField(const mxArray *s, mwIndex i, const char *name);
mxArray *mxGetCell(const mxArray *c, mwIndex i);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
double mxGetScalar(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsComplex(const mxArray *a);
void mxDestroyArray(mxArray *a);
int mxIsStruct(const mxArray *a);
int mxIsCell(const mxArray *a);
int mxIsChar(const mxArray *a);
int mxIsEmpty(const mxArray *a);
int mxIsClass(const mxArray *a, const char *cls);
char *mxArrayToString(const mxArray *a);
void mxFree(void *p);
#ifdef __cplusplus
}
#endif
#endif
