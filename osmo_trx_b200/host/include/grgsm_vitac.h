// grgsm_vitac.h — the MLSE entry points of Transceiver52M/grgsm_vitac/grgsm_vitac.h:65-82 on the GPU library.
// get_*_imp_resp() and detect_burst_*() of one burst form one GPU call (trxb200_vitac_batch runs the CIR
// search, the matched filter and the Viterbi detector together): get_*_imp_resp() runs it and keeps the
// per-thread result for `input`, detect_burst_*() re-runs it with the start the caller passes (clamped or
// not, Transceiver.cpp:631-635, ms_upper.cpp:224-225).  detect_burst_*() therefore has to follow a
// get_*_imp_resp() on the same input in the same thread, which is how every caller in the reference uses it.
#pragma once
#include <complex>
#include <cstdint>
typedef std::complex<float> gr_complex;
typedef int8_t sbit_t;

const int d_OSR(4);
void initvita();
int get_norm_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, float *corr_max, int bcc);
int get_access_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, float *corr_max, int max_delay);
int get_sch_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp);
void detect_burst_nb(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary);
void detect_burst_ab(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary);
