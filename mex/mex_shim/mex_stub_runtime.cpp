// mex_stub_runtime.cpp -- a minimal stand-in for libmx / libmex, TEST INFRASTRUCTURE ONLY.
//
// MATLAB and Octave are absent from the build image and the GPU boxes, so the MEX gateways (mex/*.cpp) could only be
// syntax-checked against the prototype shim mex.h.  This file implements the handful of mx* / mex* entry points the gateways
// use -- double / complex-double arrays, struct arrays, mexCallMATLAB("randn" / "rand") fed from a caller-supplied stream,
// mexErrMsgIdAndTxt as an exception caught at the harness boundary -- so that the gateway sources can be LINKED against
// librsp.so and EXECUTED: tests/test_mex_gateway.py builds the reference's argument structs through these functions, calls
// mexFunction, and compares what comes back with the ctypes path on the same noise.  It says nothing about MATLAB's own
// memory manager or name resolution; the genuine build is `mex -R2018a` / `mkoctfile --mex`.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "mex.h"

struct mxArray_tag {
    bool is_struct = false, is_complex = false;
    std::vector<mwSize> dims;
    std::vector<double> re;                         // real data, or interleaved (re, im) pairs when is_complex
    std::vector<std::string> field_names;           // struct arrays: fields[element][field]
    std::vector<std::vector<mxArray*>> fields;
    size_t numel() const { size_t n = 1; for (mwSize d : dims) n *= d; return n; }
};

namespace {
struct MexError : std::runtime_error {
    std::string id;
    MexError(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {}
};
std::vector<double> g_randn;                        // the stream mexCallMATLAB("randn") / ("rand") consumes
size_t g_randn_pos = 0;
std::string g_last_error, g_printed;
void (*g_at_exit)(void) = nullptr;
int g_locks = 0;

mxArray* new_numeric(const std::vector<mwSize>& dims, bool cplx) {
    mxArray* a = new mxArray_tag;
    a->dims = dims;
    a->is_complex = cplx;
    a->re.assign(a->numel() * (cplx ? 2 : 1), 0.0);
    return a;
}
}  // namespace

extern "C" {

int mxIsStruct(const mxArray* a) { return a && a->is_struct; }
int mxIsDouble(const mxArray* a) { return a && !a->is_struct; }
int mxIsComplex(const mxArray* a) { return a && a->is_complex; }
int mxIsEmpty(const mxArray* a) { return !a || a->numel() == 0; }
size_t mxGetNumberOfElements(const mxArray* a) { return a ? a->numel() : 0; }
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a ? a->dims.size() : 0; }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
mxArray* mxGetField(const mxArray* a, mwIndex i, const char* name) {
    if (!a || !a->is_struct || i >= a->fields.size()) return nullptr;
    for (size_t f = 0; f < a->field_names.size(); ++f)
        if (a->field_names[f] == name) return a->fields[i][f];
    return nullptr;
}
double mxGetScalar(const mxArray* a) { return (a && !a->re.empty()) ? a->re[0] : 0.0; }
double* mxGetDoubles(const mxArray* a) { return const_cast<double*>(a->re.data()); }
mxComplexDouble* mxGetComplexDoubles(const mxArray* a) { return reinterpret_cast<mxComplexDouble*>(const_cast<double*>(a->re.data())); }
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return new_numeric({m, n}, c == mxCOMPLEX); }
mxArray* mxCreateNumericArray(mwSize nd, const mwSize* d, mxClassID, mxComplexity c) {
    return new_numeric(std::vector<mwSize>(d, d + nd), c == mxCOMPLEX);
}
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nf, const char** names) {
    mxArray* a = new mxArray_tag;
    a->is_struct = true;
    a->dims = {m, n};
    for (int f = 0; f < nf; ++f) a->field_names.push_back(names[f]);
    a->fields.assign(m * n, std::vector<mxArray*>((size_t)nf, nullptr));
    return a;
}
mxArray* mxCreateDoubleScalar(double v) {
    mxArray* a = new_numeric({1, 1}, false);
    a->re[0] = v;
    return a;
}
void mxSetField(mxArray* a, mwIndex i, const char* name, mxArray* v) {
    for (size_t f = 0; f < a->field_names.size(); ++f)
        if (a->field_names[f] == name) { a->fields[i][f] = v; return; }
    a->field_names.push_back(name);                 // the stub lets the harness add fields on the fly
    for (auto& el : a->fields) el.push_back(nullptr);
    a->fields[i].back() = v;
}
void mxDestroyArray(mxArray* a) {
    if (!a) return;
    for (auto& el : a->fields)
        for (mxArray* f : el) mxDestroyArray(f);
    delete a;
}
void* mxMalloc(size_t n) { return std::malloc(n); }
void mxFree(void* p) { std::free(p); }

// randn(P, N) / rand(): consecutive column-major blocks of the caller's stream, exactly the order fsf:81-88 draws them in
int mexCallMATLAB(int nlhs, mxArray** plhs, int nrhs, mxArray** prhs, const char* fn) {
    const std::string f = fn;
    if (nlhs != 1 || (f != "randn" && f != "rand")) throw MexError("stub:call", "mexCallMATLAB(" + f + ") is not provided by the stub");
    mwSize m = 1, n = 1;
    if (nrhs == 1) { m = (mwSize)prhs[0]->re[0]; n = (mwSize)prhs[0]->re[1]; }
    mxArray* a = new_numeric({m, n}, false);
    if (g_randn_pos + a->re.size() > g_randn.size()) { mxDestroyArray(a); throw MexError("stub:randn", "the random stream is exhausted"); }
    std::memcpy(a->re.data(), g_randn.data() + g_randn_pos, a->re.size() * sizeof(double));
    g_randn_pos += a->re.size();
    plhs[0] = a;
    return 0;
}
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw MexError(id, buf);                        // MATLAB longjmps out of the MEX file; the stub unwinds to the harness
}
int mexPrintf(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_printed += buf;
    return n;
}
void mexLock(void) { ++g_locks; }
int mexAtExit(void (*fn)(void)) { g_at_exit = fn; return 0; }

// ------------------------------------------------------------------------------------------ harness (ctypes)
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

void stub_set_random_stream(const double* data, size_t n) { g_randn.assign(data, data + n); g_randn_pos = 0; }
size_t stub_random_consumed(void) { return g_randn_pos; }
const char* stub_last_error(void) { return g_last_error.c_str(); }
const char* stub_printed(void) { return g_printed.c_str(); }
int stub_lock_count(void) { return g_locks; }
// returns 0, or 1 with stub_last_error() = "<id>: <message>" when the gateway raised an error
int stub_call_mex(int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs) {
    g_last_error.clear();
    try {
        mexFunction(nlhs, plhs, nrhs, prhs);
        return 0;
    } catch (const MexError& e) {
        g_last_error = e.id + ": " + e.what();
        return 1;
    }
}
void stub_run_at_exit(void) { if (g_at_exit) g_at_exit(); }

}  // extern "C"
