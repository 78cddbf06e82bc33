// rsp_mex_common.hpp -- helpers shared by the two MEX gateways (struct unpacking, context cache).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"
#include "rsp.h"

namespace rspmex {

inline const mxArray* field(const mxArray* s, const char* name, const char* where) {
    const mxArray* f = (s && mxIsStruct(s)) ? mxGetField(s, 0, name) : nullptr;
    if (!f) mexErrMsgIdAndTxt("rsp:missingField", "field %s.%s is missing", where, name);
    return f;
}
inline const mxArray* field_at(const mxArray* s, size_t idx, const char* name) {
    const mxArray* f = mxGetField(s, idx, name);
    if (!f) mexErrMsgIdAndTxt("rsp:missingField", "targets(%d).%s is missing", (int)idx + 1, name);
    return f;
}
inline double scalar(const mxArray* s, const char* name, const char* where) { return mxGetScalar(field(s, name, where)); }

inline std::vector<double> reals(const mxArray* a, const char* what) {
    if (!a || !mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("rsp:type", "%s must be a real double array", what);
    const double* p = mxGetDoubles(a);
    return std::vector<double>(p, p + mxGetNumberOfElements(a));
}
inline std::vector<rsp_c128> complexes(const mxArray* a, const char* what) {
    if (!a || !mxIsDouble(a)) mexErrMsgIdAndTxt("rsp:type", "%s must be a double array", what);
    const size_t n = mxGetNumberOfElements(a);
    std::vector<rsp_c128> out(n);
    if (mxIsComplex(a)) {
        const mxComplexDouble* p = mxGetComplexDoubles(a);
        for (size_t i = 0; i < n; ++i) out[i] = rsp_c128{p[i].real, p[i].imag};
    } else {
        const double* p = mxGetDoubles(a);
        for (size_t i = 0; i < n; ++i) out[i] = rsp_c128{p[i], 0.0};
    }
    return out;
}

// One context per MATLAB process (parfor workers are separate processes => one context each),
// rebuilt when the shape or the detector parameters change; freed by mexAtExit.
struct Cache {
    rsp_ctx* ctx = nullptr;
    rsp_params prm{};
    bool waveform = false;        // rsp_set_waveform done for this context (device synthesis mode)
};
inline Cache& cache() { static Cache c; return c; }
inline void release() {
    if (cache().ctx) { rsp_destroy(cache().ctx); cache().ctx = nullptr; }
    cache().waveform = false;
}
inline void fail(rsp_ctx* ctx, int rc, const char* what) {
    std::string msg = rsp_last_error(ctx);      // copy before anything is destroyed
    char id[32];
    std::snprintf(id, sizeof id, "rsp:e%d", -rc);
    mexErrMsgIdAndTxt(id, "%s failed (%d): %s", what, rc, msg.c_str());   // longjmps out of the MEX file
}

}  // namespace rspmex
