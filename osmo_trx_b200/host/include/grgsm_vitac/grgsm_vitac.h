// grgsm_vitac.h — the MLSE entry points of Transceiver52M/grgsm_vitac/grgsm_vitac.h:65-82 on the GPU library.
//
// get_*_imp_resp(): channel-estimate search on the device (trxb200_vitac_batch), returns the burst start and fills the
// 20-tap estimate and corr_max, nothing is kept between calls.  detect_burst_*(): matched filter + Viterbi detector on the
// device (trxb200_vitac_detect_ss_batch) with the channel estimate and the start the CALLER passes - whatever the caller
// did with either in between (clamping, Transceiver.cpp:631-635 / ms_upper.cpp:224-225; pointer arithmetic,
// ms_rx_lower.cpp:177) is honoured.  Every call is a batch of one: this layer is for drop-in correctness, throughput
// comes from the batched C ABI.
//
// Memory around `input`: the reference reads input[burst_start .. burst_start + 4 N) with burst_start possibly
// negative and leaves it to the caller to own that memory.  This mirror ships kVitacAvail samples starting at `input` to
// the device and, when the caller has declared head-room with vitac_input_headroom(h), the h samples before `input`
// as well; samples it was not given read as zero (equal to the reference whenever the caller's pad is zero, as in
// ms_upper.cpp:164-171).
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>
#include "constants.h"

extern "C" {
#include <osmocom/core/bits.h>
}

#define SYNC_SEARCH_RANGE 30
const int d_OSR(4);

void initvita(); // fills the reference-symbol tables below from the GPU context (after sigProcLibSetup())

// encoded (conjugated) training sequences, as the reference exports them (grgsm_vitac.cpp:46-49; burst-gen.cpp:156-159)
extern gr_complex d_acc_training_seq[N_ACCESS_BITS];
extern gr_complex d_sch_training_seq[N_SYNC_BITS];
extern gr_complex d_norm_training_seq[TRAIN_SEQ_NUM][N_TRAIN_BITS];

int get_norm_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, float *corr_max, int bcc);
int get_access_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, float *corr_max, int max_delay);
int get_sch_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp);
int get_sch_buffer_chan_imp_resp(const gr_complex *input, gr_complex *chan_imp_resp, unsigned int len, float *corr_max);
void detect_burst_nb(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary, int ss);
void detect_burst_ab(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary, int ss);
void detect_burst_nb(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary);
void detect_burst_ab(const gr_complex *input, gr_complex *chan_imp_resp, int burst_start, sbit_t *output_binary);

// this mirror's extension: samples the caller owns BEFORE `input` (default 0) and from `input` on (default 625)
void vitac_input_headroom(int samples_before, int samples_from_input = 625);

enum class btype { NB, SCH };
struct fdata {
	btype t;
	unsigned int fn;
	int tn;
	int bcc;
	std::string fpath;
	std::vector<gr_complex> data;
	unsigned int data_start_offset;
};
