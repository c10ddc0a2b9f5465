// Host-side tap and window design for /comms/fir_designer and /comms/window_designer.
//
// The reference delegates this maths to the external Spuce library (design_fir,
// design_complex_fir, design_window, remez_estimate_*: filter/FIRDesigner.cpp:9-19,426-438,467,
// window/WindowDesigner.cpp:129), which is neither vendored in the reference tree nor present in
// this image.  These are independent implementations of the same textbook designs behind the
// same entry points; the reference pins their results only at the level of a pass/stop mask
// (filter/TestFIRDesigner.cpp:103-124), which tests/test_designers_cpu.py re-runs.
#pragma once
#include <complex>
#include <string>
#include <vector>

namespace b200c_design {

// window: "rectangular", "hann", "hamming", "blackman", "bartlett", "flattop", "kaiser" (arg = beta),
// "chebyshev" (arg = side-lobe attenuation in dB); throws std::runtime_error on an unknown name
std::vector<double> design_window(const std::string &type, size_t n, double arg);

// type: "sinc", "maxflat", "gaussian", "remez", "raised_cosine", "root_raised_cosine" (lower case);
// band: "LOW_PASS", "HIGH_PASS", "BAND_PASS", "BAND_STOP"; fl, fu normalised to the sample rate;
// alpha: excess bandwidth (cosine types) or transition width (remez); weight: remez pass/stop weight
std::vector<double> design_fir(const std::string &type, const std::string &band, size_t n, double fl, double fu, double alpha,
                               double weight);
// band: "COMPLEX_BAND_PASS", "COMPLEX_BAND_STOP"
std::vector<std::complex<double>> design_complex_fir(const std::string &type, const std::string &band, size_t n, double fl,
                                                     double fu, double alpha, double weight);

// Kaiser / Bellanger style estimates for equiripple designs (trans_bw normalised to the sample rate)
size_t remez_estimate_num_taps(double trans_bw, double pass_db, double stop_db);
double remez_estimate_weight(double pass_db, double stop_db);
double remez_estimate_bw(size_t num_taps, double pass_db, double stop_db);
double remez_estimate_atten(size_t num_taps, double trans_bw, double pass_db);

// Parks-McClellan low-pass (linear phase, n taps): pass band [0, pass_edge], stop band
// [stop_edge, 0.5], error weight 1 in the pass band and `stop_weight` in the stop band
std::vector<double> remez_lowpass(size_t n, double pass_edge, double stop_edge, double stop_weight);

} // namespace b200c_design
