// process_stage2_mtd_mex.cpp -- MEX gateway that shadows
//   [MTD_results, PC_results] = process_stage2_mtd(iq_data, angle, config)
// (/root/reference/Simulation/process_stage2_mtd.m:1).  iq_data is the already beamformed and
// range-gated cube [P, G, B] (process_stage2_mtd.m:29-30 hard-codes 332 x 3404 x 13); `angle` is
// unused by the reference (process_stage2_mtd.m:26).  The per-beam callee fun_MTD_produce is not
// shipped by the reference; the semantics implemented by rsp_stage2_mtd are specified in DESIGN.md.
//
// The context is built from `config` on the first call and rebuilt when the shape changes:
//   config.Sig_Config.{prtNum, point_prt(2:4), fs, tao(1:3), B}   (main_test_with_simulated_data.m:46-140)
//   config.mtd.beam_num                                           (process_stage2_mtd.m:15)
//   config.cfar.MTD_0v_num (optional, 0 = no zero-velocity notch) (main_test_with_simulated_data.m:106,124)
//   config.mtd.win (optional real [P] Doppler window; absent = rectangular)
#include "rsp_mex_common.hpp"

using namespace rspmex;

namespace {

const double kPi = 3.14159265358979323846;

// debug_simulated_data_processing_v2.m:309-317: t = -tau/2 : ts : tau/2 - ts;
// pulse1 = sin(2*pi*t + pi/2), pulse2/3 = exp(j*2*pi*(0.5*K*t.^2)), K2 = -B/tau2, K3 = +B/tau3.
std::vector<rsp_c128> reference_pulse(int seg, double tau, double fs, double bw) {
    const double ts = 1.0 / fs;
    const int n = (int)std::llround(tau / ts);
    std::vector<rsp_c128> p((size_t)n);
    const double K = seg == 1 ? -bw / tau : bw / tau;
    for (int i = 0; i < n; ++i) {
        const double t = -tau / 2 + ts * i;
        if (seg == 0) p[(size_t)i] = rsp_c128{std::sin(2 * kPi * t + kPi / 2), 0.0};
        else { const double ph = 2 * kPi * (0.5 * K * t * t); p[(size_t)i] = rsp_c128{std::cos(ph), std::sin(ph)}; }
    }
    return p;
}

void ensure_context(const mxArray* config, const mwSize* d) {
    const mxArray* sc = field(config, "Sig_Config", "config");
    const mxArray* mtd = field(config, "mtd", "config");
    const std::vector<double> pts = reals(field(sc, "point_prt", "config.Sig_Config"), "config.Sig_Config.point_prt");
    const std::vector<double> tao = reals(field(sc, "tao", "config.Sig_Config"), "config.Sig_Config.tao");
    if (pts.size() < 4 || tao.size() < 3)
        mexErrMsgIdAndTxt("rsp:config", "config.Sig_Config.point_prt needs 4 entries and .tao 3");
    rsp_params p{};
    p.abi_version = RSP_ABI_VERSION;
    p.n_channels = 1;
    p.n_beams = (int32_t)scalar(mtd, "beam_num", "config.mtd");
    p.n_pulses = (int32_t)scalar(sc, "prtNum", "config.Sig_Config");
    int32_t G = 0;
    for (int s = 0; s < 3; ++s) { p.n_gates[s] = (int32_t)pts[(size_t)s + 1]; p.seg_start[s] = 1; G += p.n_gates[s]; }
    p.n_samples = G > 64 ? G : 64;
    p.t_cfar = 8.0;                                   // the detector is not run on this path
    p.guard_r = p.guard_v = p.ref_r = p.ref_v = 1;
    p.max_detections = 16;
    if (d[0] != (mwSize)p.n_pulses || d[1] != (mwSize)G || d[2] != (mwSize)p.n_beams)
        mexErrMsgIdAndTxt("rsp:shape", "iq_data is %d x %d x %d, config describes %d x %d x %d", (int)d[0], (int)d[1],
                          (int)d[2], p.n_pulses, (int)G, p.n_beams);
    Cache& c = cache();
    if (c.ctx && std::memcmp(&c.prm, &p, sizeof p) == 0) return;
    release();

    const double fs = scalar(sc, "fs", "config.Sig_Config"), bw = scalar(sc, "B", "config.Sig_Config");
    std::vector<rsp_c128> pulse[3];
    rsp_stage2_config s2{};
    for (int s = 0; s < 3; ++s) {
        pulse[s] = reference_pulse(s, tao[(size_t)s], fs, bw);
        s2.pulse[s] = pulse[s].data();
        s2.n_pulse[s] = (int32_t)pulse[s].size();
    }
    std::vector<double> win;
    if (const mxArray* w = mxGetField(mtd, 0, "win")) {
        win = reals(w, "config.mtd.win");
        if (win.size() != (size_t)p.n_pulses) mexErrMsgIdAndTxt("rsp:config", "config.mtd.win must have prtNum entries");
        s2.mtd_win = win.data();
    }
    const mxArray* cfar = mxIsStruct(config) ? mxGetField(config, 0, "cfar") : nullptr;
    const mxArray* notch = cfar && mxIsStruct(cfar) ? mxGetField(cfar, 0, "MTD_0v_num") : nullptr;
    s2.zero_vel_bins = notch ? (int32_t)mxGetScalar(notch) : 0;

    rsp_ctx* ctx = nullptr;
    int rc = rsp_create(&p, &ctx);
    if (rc) fail(ctx, rc, "rsp_create");
    rc = rsp_stage2_configure(ctx, &s2);
    if (rc) {
        std::string msg = rsp_last_error(ctx);
        rsp_destroy(ctx);
        mexErrMsgIdAndTxt("rsp:config", "rsp_stage2_configure failed (%d): %s", rc, msg.c_str());
    }
    c.ctx = ctx;
    c.prm = p;
    static bool hooked = false;
    if (!hooked) { mexAtExit(release); mexLock(); hooked = true; }
}

}  // namespace

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 3) mexErrMsgIdAndTxt("rsp:nargin", "process_stage2_mtd(iq_data, angle, config)");
    if (nlhs > 2) mexErrMsgIdAndTxt("rsp:nargout", "process_stage2_mtd returns [MTD_results, PC_results]");
    const mxArray* iq = prhs[0];
    if (!mxIsDouble(iq) || !mxIsComplex(iq) || mxGetNumberOfDimensions(iq) != 3)
        mexErrMsgIdAndTxt("rsp:type", "iq_data must be a complex double P x G x B array");
    const mwSize* d = mxGetDimensions(iq);
    ensure_context(prhs[2], d);
    mwSize dims[3] = {d[0], d[1], d[2]};
    plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxCOMPLEX);
    mxArray* pc = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxCOMPLEX);
    const int rc = rsp_stage2_mtd(cache().ctx, mxGetComplexDoubles(iq), RSP_C128,
                                  reinterpret_cast<rsp_c128*>(mxGetComplexDoubles(plhs[0])),
                                  reinterpret_cast<rsp_c128*>(mxGetComplexDoubles(pc)));
    if (rc) { mxDestroyArray(pc); fail(cache().ctx, rc, "rsp_stage2_mtd"); }
    if (nlhs > 1) plhs[1] = pc; else mxDestroyArray(pc);
}
