// process_stage2_mtd_mex.cpp -- MEX gateway that shadows
//   [MTD_results, PC_results] = process_stage2_mtd(iq_data, angle, config)
// (/root/reference/Simulation/process_stage2_mtd.m:1).  iq_data is the already beamformed and
// range-gated cube [P, G, B] (process_stage2_mtd.m:29-30 hard-codes 332 x 3404 x 13); `angle` is
// unused by the reference (process_stage2_mtd.m:26).  The per-beam callee fun_MTD_produce is not
// shipped by the reference; the semantics implemented by rsp_stage2_mtd are specified in DESIGN.md.
#include "rsp_mex_common.hpp"

using namespace rspmex;

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 3) mexErrMsgIdAndTxt("rsp:nargin", "process_stage2_mtd(iq_data, angle, config)");
    if (nlhs > 2) mexErrMsgIdAndTxt("rsp:nargout", "process_stage2_mtd returns [MTD_results, PC_results]");
    const mxArray* iq = prhs[0];
    if (!mxIsDouble(iq) || !mxIsComplex(iq) || mxGetNumberOfDimensions(iq) != 3)
        mexErrMsgIdAndTxt("rsp:type", "iq_data must be a complex double P x G x B array");
    if (!cache().ctx)
        mexErrMsgIdAndTxt("rsp:notReady", "call rsp_stage2_setup(config) (see INTEGRATION.md) before process_stage2_mtd");
    const mwSize* d = mxGetDimensions(iq);
    const rsp_params& p = cache().prm;
    const size_t G = (size_t)p.n_gates[0] + p.n_gates[1] + p.n_gates[2];
    if (d[0] != (mwSize)p.n_pulses || d[1] != (mwSize)G || d[2] != (mwSize)p.n_beams)
        mexErrMsgIdAndTxt("rsp:shape", "iq_data is %d x %d x %d, context expects %d x %d x %d", (int)d[0], (int)d[1], (int)d[2],
                          p.n_pulses, (int)G, p.n_beams);
    mwSize dims[3] = {d[0], d[1], d[2]};
    plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxCOMPLEX);
    mxArray* pc = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxCOMPLEX);
    const int rc = rsp_stage2_mtd(cache().ctx, mxGetComplexDoubles(iq), RSP_C128,
                                  reinterpret_cast<rsp_c128*>(mxGetComplexDoubles(plhs[0])),
                                  reinterpret_cast<rsp_c128*>(mxGetComplexDoubles(pc)));
    if (rc) { mxDestroyArray(pc); fail(cache().ctx, rc, "rsp_stage2_mtd"); }
    if (nlhs > 1) plhs[1] = pc; else mxDestroyArray(pc);
}
