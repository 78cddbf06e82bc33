// fun_process_single_frame_mex.cpp -- MEX gateway that shadows
//   final_targets = fun_process_single_frame(targets, config, cfar_params, cluster_params, precomputed_data, frame_idx)
// (/root/reference/Simulation/fun_process_single_frame.m:13).  Build it as
// fun_process_single_frame.mexa64 / .mex and put its directory ahead of Simulation/ on the path: the
// drivers (main_simulate_echoes_with_array_v8*.m:232) then call the B200 chain with no change.
//
// What runs where:
//   S4   echo synthesis (fsf:47-77) in this file, straight loops over the small target list;
//   S4.1 noise (fsf:81-88) by calling MATLAB's own randn through mexCallMATLAB, in the reference's
//        order (per channel: I then Q), so that with the same rng seed the cube is the one the .m
//        file would have built;
//   S5..S11 in librsp.so: rsp_process_frame on the MATLAB-ordered [P,N,C] complex-double cube.
#include "rsp_mex_common.hpp"

using namespace rspmex;

static void fill_params(rsp_params& p, const mxArray* config, const mxArray* cfar, const mxArray* pd) {
    const mxArray* sc = field(config, "Sig_Config", "config");
    std::memset(&p, 0, sizeof p);
    p.abi_version = RSP_ABI_VERSION;
    p.n_channels = (int)scalar(sc, "channel_num", "config.Sig_Config");          // fsf:47
    p.n_beams = (int)scalar(sc, "beam_num", "config.Sig_Config");                // fsf:92
    p.n_pulses = (int)scalar(sc, "prtNum", "config.Sig_Config");                 // fsf:47
    p.n_samples = (int)scalar(sc, "point_PRT", "config.Sig_Config");             // fsf:17
    p.seg_start[0] = (int)scalar(pd, "seg_start_narrow", "precomputed_data");    // fsf:34-36
    p.seg_start[1] = (int)scalar(pd, "seg_start_medium", "precomputed_data");
    p.seg_start[2] = (int)scalar(pd, "seg_start_long", "precomputed_data");
    p.n_gates[0] = (int)scalar(pd, "N_gate_narrow", "precomputed_data");         // fsf:30-32
    p.n_gates[1] = (int)scalar(pd, "N_gate_medium", "precomputed_data");
    p.n_gates[2] = (int)scalar(pd, "N_gate_long", "precomputed_data");
    p.fir_delay = (int)scalar(pd, "fir_delay", "precomputed_data");              // fsf:25
    p.t_cfar = (float)scalar(cfar, "T_CFAR", "cfar_params");                     // fsf:177-179 (method is ignored, fsf:199)
    p.guard_r = (int)scalar(cfar, "guardCells_R", "cfar_params");
    p.guard_v = (int)scalar(cfar, "guardCells_V", "cfar_params");
    p.ref_r = (int)scalar(cfar, "refCells_R", "cfar_params");
    p.ref_v = (int)scalar(cfar, "refCells_V", "cfar_params");
    p.max_detections = 65536;
    p.monopulse_complex = 0;
    p.device = 0;
}

static std::vector<rsp_c128> taps_of(const mxArray* pd, const char* win_name, const char* fft_name) {
    if (const mxArray* w = mxGetField(pd, 0, win_name)) return complexes(w, win_name);      // v8_3:146,148
    // only the spectrum is present (the field fsf:26-27 reads): taps = leading part of its inverse DFT
    std::vector<rsp_c128> H = complexes(field(pd, fft_name, "precomputed_data"), fft_name);
    const size_t n = H.size();
    std::vector<rsp_c128> h(n);
    double peak = 0;
    for (size_t i = 0; i < n; ++i) {
        double re = 0, im = 0;
        for (size_t k = 0; k < n; ++k) {
            const double a = 2.0 * 3.14159265358979323846 * (double)((i * k) % n) / (double)n;
            re += H[k].re * std::cos(a) - H[k].im * std::sin(a);
            im += H[k].re * std::sin(a) + H[k].im * std::cos(a);
        }
        h[i] = rsp_c128{re / n, im / n};
        peak = std::fmax(peak, std::hypot(h[i].re, h[i].im));
    }
    size_t last = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::hypot(h[i].re, h[i].im) > 1e-9 * peak) last = i;
    h.resize(last + 1);
    return h;
}

static void ensure_context(const mxArray* config, const mxArray* cfar, const mxArray* pd) {
    rsp_params p;
    fill_params(p, config, cfar, pd);
    Cache& c = cache();
    if (c.ctx && std::memcmp(&p, &c.prm, sizeof p) == 0) return;
    release();
    int rc = rsp_create(&p, &c.ctx);
    if (rc) fail(nullptr, rc, "rsp_create");
    c.prm = p;
    static bool hooked = false;
    if (!hooked) { mexAtExit(release); mexLock(); hooked = true; }
    // constants (precomputed_data, fsf:21-43)
    const mxArray* Wm = field(pd, "DBF_coeffs_data_C", "precomputed_data");          // MATLAB [B,C] column-major
    std::vector<rsp_c128> Wcm = complexes(Wm, "DBF_coeffs_data_C"), W(Wcm.size());
    const int B = p.n_beams, C = p.n_channels;
    if ((int)Wcm.size() != B * C) mexErrMsgIdAndTxt("rsp:shape", "DBF_coeffs_data_C must be beam_num x channel_num");
    for (int b = 0; b < B; ++b)
        for (int ch = 0; ch < C; ++ch) W[(size_t)b * C + ch] = Wcm[(size_t)ch * B + b];   // -> row-major [B][C]
    std::vector<double> fir = reals(field(pd, "MF_narrow", "precomputed_data"), "MF_narrow");
    std::vector<rsp_c128> mfm = taps_of(pd, "MF_medium_win", "MF_medium_fft"), mfl = taps_of(pd, "MF_long_win", "MF_long_fft");
    std::vector<double> win = reals(field(pd, "MTD_win", "precomputed_data"), "MTD_win");
    std::vector<double> ra = reals(field(pd, "range_axis", "precomputed_data"), "range_axis");
    std::vector<double> va = reals(field(pd, "velocity_axis", "precomputed_data"), "velocity_axis");
    std::vector<double> ang = reals(field(pd, "beam_angles_deg", "precomputed_data"), "beam_angles_deg");
    std::vector<double> ks = reals(field(pd, "k_slopes_LUT", "precomputed_data"), "k_slopes_LUT");
    rsp_constants k;
    std::memset(&k, 0, sizeof k);
    k.dbf_weights = W.data();
    k.fir = fir.data(); k.n_fir = (int)fir.size();
    k.mf_medium = mfm.data(); k.n_mf_medium = (int)mfm.size();
    k.mf_long = mfl.data(); k.n_mf_long = (int)mfl.size();
    k.mtd_win = win.data(); k.range_axis = ra.data(); k.velocity_axis = va.data();
    k.delta_r = scalar(pd, "deltaR", "precomputed_data"); k.delta_v = scalar(pd, "deltaV", "precomputed_data");
    k.beam_angles_deg = ang.data(); k.k_slopes = ks.data();
    rc = rsp_upload_constants(c.ctx, &k);
    if (rc) fail(c.ctx, rc, "rsp_upload_constants");
}

// Device synthesis mode: S4 + S4.1 run on the GPU (rsp_process_targets), so no cube is built in MATLAB memory
// and only the target list crosses the bus (tens of microseconds per frame instead of seconds).  The noise
// then comes from the library's Philox generator, seeded with one draw of MATLAB's global stream per frame
// (so rng(seed) still makes a run reproducible, but the samples differ from the .m file's randn cube).
// Enabled by config.rsp_device_synthesis = true or the environment variable RSP_MEX_DEVICE_SYNTH=1.
static bool device_synthesis(const mxArray* config) {
    const mxArray* f = mxIsStruct(config) ? mxGetField(config, 0, "rsp_device_synthesis") : nullptr;
    if (f && mxGetNumberOfElements(f) > 0) return mxGetScalar(f) != 0.0;
    const char* e = std::getenv("RSP_MEX_DEVICE_SYNTH");
    return e && std::atoi(e) != 0;
}

static long matlab_round(double x) { return x >= 0 ? (long)std::floor(x + 0.5) : -(long)std::floor(-x + 0.5); }

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 5) mexErrMsgIdAndTxt("rsp:nargin", "fun_process_single_frame needs 5 or 6 inputs");
    if (nlhs > 1) mexErrMsgIdAndTxt("rsp:nargout", "fun_process_single_frame returns one output");
    const mxArray *targets = prhs[0], *config = prhs[1], *cfar = prhs[2], *clus = prhs[3], *pd = prhs[4];
    ensure_context(config, cfar, pd);
    const rsp_params& p = cache().prm;
    const mxArray* sc = field(config, "Sig_Config", "config");
    const size_t P = p.n_pulses, N = p.n_samples, C = p.n_channels;
    const double c0 = scalar(sc, "c", "config.Sig_Config"), fs = scalar(sc, "fs", "config.Sig_Config");
    const double lam = scalar(sc, "wavelength", "config.Sig_Config"), prt = scalar(sc, "prt", "config.Sig_Config");
    const double d_el = scalar(field(config, "Array", "config"), "element_spacing", "config.Array");
    const std::vector<rsp_c128> tx = complexes(field(pd, "tx_pulse", "precomputed_data"), "tx_pulse");
    const double p_sig = scalar(pd, "P_signal_unscaled", "precomputed_data");

    rsp_cluster_params cp{scalar(clus, "max_range_sep", "cluster_params"), scalar(clus, "max_vel_sep", "cluster_params"),
                          scalar(clus, "max_angle_sep", "cluster_params")};
    std::vector<rsp_target> fin(4096);
    int32_t nf = 0;
    const size_t K = mxIsStruct(targets) ? mxGetNumberOfElements(targets) : 0;

    if (device_synthesis(config)) {
        if (!cache().waveform) {
            rsp_waveform w{tx.data(), c0, fs, lam, prt, d_el, p_sig};
            if ((size_t)tx.size() != N) mexErrMsgIdAndTxt("rsp:shape", "tx_pulse must have point_PRT entries");
            const int rcw = rsp_set_waveform(cache().ctx, &w);
            if (rcw) fail(cache().ctx, rcw, "rsp_set_waveform");
            cache().waveform = true;
        }
        std::vector<rsp_target_in> tin(K);
        for (size_t k = 0; k < K; ++k)
            tin[k] = rsp_target_in{mxGetScalar(field_at(targets, k, "Range")), mxGetScalar(field_at(targets, k, "Velocity")),
                                   mxGetScalar(field_at(targets, k, "ElevationAngle")), mxGetScalar(field_at(targets, k, "SNR_dB"))};
        mxArray* u = nullptr;
        mexCallMATLAB(1, &u, 0, nullptr, "rand");                       // one draw of the global stream per frame
        const uint64_t seed = (uint64_t)std::ldexp(mxGetScalar(u), 53);
        mxDestroyArray(u);
        const int rc = rsp_process_targets(cache().ctx, tin.data(), (int32_t)K, 1.0 /* P_noise, fsf:83 */, seed, &cp, fin.data(),
                                           (int32_t)fin.size(), &nf, nullptr, 0, nullptr);
        if (rc) fail(cache().ctx, rc, "rsp_process_targets");
    } else {
    // S4 (fsf:47-77): raw(p, n, c) in MATLAB order, element index (c*N + n)*P + p
    mwSize dims[3] = {P, N, C};
    mxArray* cube = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxCOMPLEX);
    mxComplexDouble* raw = mxGetComplexDoubles(cube);
    const double two_pi = 6.28318530717958647692;
    for (size_t k = 0; k < K; ++k) {
        const double R = mxGetScalar(field_at(targets, k, "Range")), V = mxGetScalar(field_at(targets, k, "Velocity"));
        const double El = mxGetScalar(field_at(targets, k, "ElevationAngle")), snr = mxGetScalar(field_at(targets, k, "SNR_dB"));
        const long d = matlab_round((2.0 * R / c0) / (1.0 / fs));                          // fsf:18,55-56: round(delay / ts)
        if (!(d > 0 && d < (long)N)) continue;                                            // fsf:66
        const size_t len = std::min(tx.size(), N - (size_t)d);                            // fsf:67
        const double amp = std::sqrt(std::pow(10.0, snr / 10.0) / p_sig);                 // fsf:61-63
        const double fd = 2.0 * V / lam;                                                  // fsf:57
        const double dphi = two_pi * d_el * std::sin(El * 3.14159265358979323846 / 180.0) / lam;   // fsf:163-169
        for (size_t ch = 0; ch < C; ++ch) {
            const double cr = std::cos(ch * dphi), ci = std::sin(ch * dphi);              // fsf:72
            for (size_t i = 0; i < len; ++i) {
                if (tx[i].re == 0.0 && tx[i].im == 0.0) continue;
                const double br = amp * (tx[i].re * cr - tx[i].im * ci), bi = amp * (tx[i].re * ci + tx[i].im * cr);
                mxComplexDouble* col = raw + (ch * N + (size_t)d + i) * P;
                for (size_t m = 0; m < P; ++m) {
                    const double a = two_pi * fd * (double)m * prt;                      // fsf:58
                    const double dr = std::cos(a), di = std::sin(a);
                    col[m].real += br * dr - bi * di;
                    col[m].imag += br * di + bi * dr;
                }
            }
        }
    }
    // S4.1 (fsf:81-88): MATLAB's own randn, channel by channel, I then Q, scaled by sqrt(1/2)
    mxArray* sz = mxCreateDoubleMatrix(1, 2, mxREAL);
    mxGetDoubles(sz)[0] = (double)P; mxGetDoubles(sz)[1] = (double)N;
    const double sigma = std::sqrt(0.5);
    for (size_t ch = 0; ch < C; ++ch) {
        mxArray *I = nullptr, *Q = nullptr;
        mexCallMATLAB(1, &I, 1, &sz, "randn");
        mexCallMATLAB(1, &Q, 1, &sz, "randn");
        const double *pi_ = mxGetDoubles(I), *pq = mxGetDoubles(Q);
        mxComplexDouble* dst = raw + ch * N * P;
        for (size_t i = 0; i < N * P; ++i) { dst[i].real += sigma * pi_[i]; dst[i].imag += sigma * pq[i]; }
        mxDestroyArray(I);
        mxDestroyArray(Q);
    }
    mxDestroyArray(sz);

    // S5..S11 on the GPU
    const int rc = rsp_process_frame(cache().ctx, raw, RSP_LAYOUT_MATLAB, RSP_C128, RSP_MEM_HOST, &cp, fin.data(), (int32_t)fin.size(), &nf);
    mxDestroyArray(cube);
    if (rc) fail(cache().ctx, rc, "rsp_process_frame");
    }

    if (nf == 0) {                                  // fsf:229-232,305-308,358-361: [] when nothing is detected
        plhs[0] = mxCreateDoubleMatrix(0, 0, mxREAL);
        return;
    }
    const char* names[4] = {"Range", "Velocity", "Angle", "Power"};     // fsf:393
    plhs[0] = mxCreateStructMatrix((mwSize)nf, 1, 4, names);            // n x 1, like repmat(struct(...), n, 1)
    for (int32_t i = 0; i < nf; ++i) {
        mxSetField(plhs[0], i, "Range", mxCreateDoubleScalar(fin[i].range));
        mxSetField(plhs[0], i, "Velocity", mxCreateDoubleScalar(fin[i].velocity));
        mxSetField(plhs[0], i, "Angle", mxCreateDoubleScalar(fin[i].angle));
        mxSetField(plhs[0], i, "Power", mxCreateDoubleScalar(fin[i].power));
    }
    if (nrhs >= 6) mexPrintf("  > frame %d done, %d unique targets.\n", (int)mxGetScalar(prhs[5]), (int)nf);   // fsf:156
}
