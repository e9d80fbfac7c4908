/* oracle_sched.c — TEST INFRASTRUCTURE ONLY (see oracle_trx.h).
 * CPU restatement of the burst-type scheduler around the hot path:
 * Transceiver::expectedCorrType (Transceiver52M/Transceiver.cpp:513-601) and the search-window choice of
 * pullRadioVector (:757-758).  Pinned bit-for-bit against the reference's own function (oracle/_ref, cut out of
 * Transceiver.cpp at build time by gen_ref_sched.py) and against tests/golden/sched_fixture.npz. */
#include <stdint.h>

/* ChannelCombination, Transceiver.h:131-148 */
enum { CC_FILL, CC_I, CC_II, CC_III, CC_IV, CC_V, CC_VI, CC_VII, CC_VIII, CC_IX, CC_X, CC_XI, CC_XII, CC_XIII, CC_NONE, CC_LOOPBACK };
/* CorrType, sigProcLib.h:29-37 */
enum { T_OFF = 0, T_TSC = 1, T_EXT_RACH = 2, T_RACH = 3, T_SCH = 4, T_EDGE = 5, T_IDLE = 6 };

/* half-rate sub-slot of a 26-multiframe position (:516) */
static int tchh_subslot(unsigned fn26)
{
	if (fn26 < 12) return fn26 & 1;  /* 0,1,0,1,... */
	if (fn26 == 12) return 0;	  /* 0,0 at 12,13 */
	if (fn26 < 25) return (fn26 & 1) ? 0 : 1; /* 13:0 14:1 15:0 ... 24:1 */
	return 1;			  /* 25 */
}

/* SDCCH/4 and SDCCH/8 sub-slot per 102-multiframe position (:517-520; 3GPP TS 45.002 uplink mapping): irregular
 * enough that the restatement keeps them as data, as (value, run length) pairs */
static void expand(const unsigned char *rl, int nrl, int *out)
{
	int k = 0;
	for (int i = 0; i < nrl; i++)
		for (int c = 0; c < rl[2 * i + 1]; c++) out[k++] = rl[2 * i];
}

static const unsigned char SD4_RL[] = { 3,4, 0,2, 2,4, 3,4, 0,27, 1,4, 0,2, 2,4, 3,4, 0,6, 1,4, 0,27, 1,4, 0,2, 2,4 };
static const unsigned char SD8_RL[] = { 5,4, 6,4, 7,4, 0,7, 1,4, 2,4, 3,4, 4,4, 5,4, 6,4, 7,4, 0,4, 1,4, 2,4, 3,4, 0,7, 1,4, 2,4, 3,4, 4,4, 5,4, 6,4, 7,4, 4,4 };

int orc_expected_corr_type(const uint8_t *chan_type, const uint8_t *handover, int ext_rach, int egprs, const uint32_t *fn,
			      const uint8_t *tn, int n, uint8_t *out)
{
	int sd4[102], sd8[102];
	expand(SD4_RL, (int)(sizeof(SD4_RL) / 2), sd4);
	expand(SD8_RL, (int)(sizeof(SD8_RL) / 2), sd8);
	const int rach = ext_rach ? T_EXT_RACH : T_RACH;
	for (int i = 0; i < n; i++) {
		const unsigned t = tn[i] & 7, f = fn[i];
		const unsigned ho = handover[t];
		int r = T_OFF;
		switch (chan_type[t]) {
		case CC_NONE: r = T_OFF; break;
		case CC_FILL: r = T_IDLE; break;
		case CC_I: r = (ho & 1) ? T_RACH : T_TSC; break;
		case CC_II:
			if (tchh_subslot(f % 26) == 1) r = T_IDLE;
			else r = (ho & 1) ? T_RACH : T_TSC;
			break;
		case CC_III: r = ((ho >> tchh_subslot(f % 26)) & 1) ? T_RACH : T_TSC; break;
		case CC_IV:
		case CC_VI: r = rach; break;
		case CC_V: {
			const unsigned m = f % 51;
			if ((m >= 14 && m <= 36) || m == 4 || m == 5 || m == 45 || m == 46) r = rach;
			else r = ((ho >> sd4[f % 102]) & 1) ? T_RACH : T_TSC;
			break;
		}
		case CC_VII: {
			const unsigned m = f % 51;
			if (m >= 12 && m <= 14) r = T_IDLE;
			else r = ((ho >> sd8[f % 102]) & 1) ? T_RACH : T_TSC;
			break;
		}
		case CC_XIII: {
			const unsigned m = f % 52;
			if (m == 12 || m == 38) r = T_RACH; /* PTCCH/U: always the 8-bit access burst */
			else if (m == 25 || m == 51) r = T_IDLE;
			else r = egprs ? T_EDGE : T_TSC;
			break;
		}
		case CC_LOOPBACK: {
			const unsigned m = f % 51;
			r = (m >= 48) ? T_IDLE : T_TSC;
			break;
		}
		default: r = T_OFF; break;
		}
		out[i] = (uint8_t)r;
	}
	return 0;
}

/* pullRadioVector :757-758 */
void orc_sched_max_toa(const uint8_t *type, int n, int max_toa_nb, int max_toa_ab, uint16_t *out)
{
	for (int i = 0; i < n; i++)
		out[i] = (uint16_t)((type[i] == T_RACH || type[i] == T_EXT_RACH) ? max_toa_ab : max_toa_nb);
}
