// constants.h — the GSM burst-structure constants that callers of the MLSE interface use (the reference keeps them in
// Transceiver52M/grgsm_vitac/constants.h:26-125; utils/va-test/burst-gen.cpp:156-159,293 sizes its arrays with them).
// The numbers are 3GPP TS 45.002 burst formats; the names are the interface.
#pragma once
#include <complex>

#ifndef gr_complex
#define gr_complex std::complex<float> /* constants.h:26 */
#endif

// symbol clock: 13 MHz / 48
#define GSM_SYMBOL_RATE (1625000.0 / 6.0)
#define GSM_SYMBOL_PERIOD (1.0 / GSM_SYMBOL_RATE)

// normal burst: 3 tail | 57 data | 1 stealing | 26 training | 1 stealing | 57 data | 3 tail | 8.25 guard
#define TAIL_BITS 3
#define GUARD_BITS 8
#define GUARD_FRACTIONAL 0.25
#define GUARD_PERIOD GUARD_BITS + GUARD_FRACTIONAL
#define DATA_BITS 57
#define STEALING_BIT 1
#define N_TRAIN_BITS 26
#define USEFUL_BITS 142
#define BURST_SIZE (USEFUL_BITS + 2 * TAIL_BITS)
#define PROCESSED_CHUNK BURST_SIZE + 2 * GUARD_PERIOD
#define TS_BITS (TAIL_BITS + USEFUL_BITS + TAIL_BITS + GUARD_BITS)
#define TS_PER_FRAME 8
#define FRAME_BITS (TS_PER_FRAME * TS_BITS + 2)
// the channel estimate skips the first five training bits (grgsm_vitac.cpp:265-274)
#define TRAIN_BEGINNING 5
#define TRAIN_POS (TAIL_BITS + (DATA_BITS + STEALING_BIT) + TRAIN_BEGINNING)
#define SAFETY_MARGIN 6

// synchronisation burst (64-bit extended training sequence after 3 tail + 39 data bits), frequency-correction burst,
// access burst (8 extended tail + 41 sync + 36 data + 3 tail)
#define N_SYNC_BITS 64
#define SCH_DATA_LEN 39
#define SYNC_POS (TAIL_BITS + SCH_DATA_LEN)
#define FCCH_BITS USEFUL_BITS
#define FCCH_POS TAIL_BITS
#define FCCH_HITS_NEEDED (USEFUL_BITS - 4)
#define FCCH_MAX_MISSES 1
#define FCCH_MAX_FREQ_OFFSET 100
#define N_ACCESS_BITS 41
#define ACCESS_BURST_SIZE 88
#define MAX_SCH_ERRORS 10

// channel impulse response length in symbols (x d_OSR samples)
#define CHAN_IMP_RESP_LENGTH 5

typedef enum { empty, fcch_burst, sch_burst, normal_burst, rach_burst, dummy, dummy_or_normal, normal_or_noise } burst_type;
typedef enum { unknown, multiframe_26, multiframe_51 } multiframe_type;

// training sequence codes: eight TSCs plus the dummy burst's sequence
#define TSC0 0
#define TSC1 1
#define TSC2 2
#define TSC3 3
#define TSC4 4
#define TSC5 5
#define TSC6 6
#define TSC7 7
#define TS_DUMMY 8
#define TRAIN_SEQ_NUM 9

#define TIMESLOT0 0
#define TIMESLOT1 1
#define TIMESLOT2 2
#define TIMESLOT3 3
#define TIMESLOT4 4
#define TIMESLOT5 5
#define TIMESLOT6 6
#define TIMESLOT7 7
