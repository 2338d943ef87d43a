// aisgpu.cu -- host side of the C ABI declared in include/aisgpu.h.
//
// Mirrors, for a batch of streams, what AIS::ModelFrontend::buildModel + ModelDefault/Standard/Base::buildModel
// wire up for one stream (reference Source/DSP/Model.cpp:27-356, 419-438, 484-577): the rate -> chain table, the
// filter parameters, and the per-submit launch sequence of the kernels in aisgpu_kernels.cuh.  The only
// arithmetic done on the host is the libm-dependent constant tables (Rotate step, FFT twiddles, CGF phasor
// steps, Model.cpp:31, FFT.h:81-83, DSP.cpp:457-458) and the per-frame tail of AIS::Decoder::processData
// (dB level, validate, buildNMEA: AIS.cpp:66-96, Message.cpp:398-413, 569-686).  No CPU fallback exists.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <string>
#include <vector>

#include "../../include/aisgpu.h"
#include "params.h"

using namespace aisgpu;

namespace {

thread_local std::string g_create_error;

const float PI_F = 3.14159265358979323846f; // Library/Common.h:318

float2 polar1(float theta) { // std::polar(1.0f, theta) through sincosf, as the reference build resolves it
	float s, c;
	sincosf(theta, &s, &c);
	return make_float2(1.0f * c, 1.0f * s);
}

const float H_TAPS_RECEIVER[37] = { // DSP/Filters.h:24-33
	0.00119025f, -0.00148464f, -0.00282428f, -0.00200561f, -0.00068852f, 0.00343044f, 0.00902093f, 0.01367867f,
	0.01147965f, 0.0027259f, -0.01766614f, -0.04244429f, -0.0577468f, -0.05245161f, -0.01072754f, 0.0732564f,
	0.17643278f, 0.25582214f, 0.28200453f, 0.25582214f, 0.17643278f, 0.0732564f, -0.01072754f, -0.05245161f,
	-0.0577468f, -0.04244429f, -0.01766614f, 0.0027259f, 0.01147965f, 0.01367867f, 0.00902093f, 0.00343044f,
	-0.00068852f, -0.00200561f, -0.00282428f, -0.00148464f, 0.00119025f };
const float H_TAPS_COHERENT[17] = { // DSP/Filters.h:35-41
	2.06995719e-06f, 3.18610148e-05f, 3.40605309e-04f, 2.52892989e-03f, 1.30411453e-02f, 4.67076746e-02f,
	1.16186141e-01f, 2.00730781e-01f, 2.40861391e-01f, 2.00730781e-01f, 1.16186141e-01f, 4.67076746e-02f,
	1.30411453e-02f, 2.52892989e-03f, 3.40605309e-04f, 3.18610148e-05f, 2.06995719e-06f };
const float H_TAPS_BH28_3[26] = { // DSP/Filters.h:43-53
	6.32542387e-05f, -2.90015252e-04f, -1.54206250e-03f, -1.64972455e-03f, 3.12793899e-03f, 1.09494413e-02f,
	9.04975801e-03f, -1.43685846e-02f, -4.45615933e-02f, -3.44883647e-02f, 5.53474269e-02f, 2.01827915e-01f,
	3.16534610e-01f, 3.16534610e-01f, 2.01827915e-01f, 5.53474269e-02f, -3.44883647e-02f, -4.45615933e-02f,
	-1.43685846e-02f, 9.04975801e-03f, 1.09494413e-02f, 3.12793899e-03f, -1.64972455e-03f, -1.54206250e-03f,
	-2.90015252e-04f, 6.32542387e-05f };
const float H_PS_COS[8] = { 9.9518472640441780e-01f, 9.5694033335306883e-01f, 8.8192125790916542e-01f, 7.7301044123076901e-01f,
							6.3439326515712957e-01f, 4.7139671032286945e-01f, 2.9028464326824349e-01f, 9.8017099547459546e-02f }; // Demod.h:29-31
const float H_PS_SIN[8] = { 9.8017143048367339e-02f, 2.9028468509743588e-01f, 4.7139674887287397e-01f, 6.3439329894649099e-01f,
							7.7301046896098113e-01f, 8.8192127851457169e-01f, 9.5694034604181499e-01f, 9.9518473068888236e-01f };

constexpr int HC = 1024; // room in front of new 48 kHz samples: unconsumed CGF samples (<512), FM/FIR history (37), or the V2 engine's
                         // block awaiting its lookahead plus a partial block (<1024)
constexpr int V2_BLK = 512; // V2::BLOCK_SIZE (V2Engine.h:30)
constexpr int HD = 48;  // ModelChallenger: derotated samples kept in front of the new ones (FM needs 1, FIR37 36, a partial group 4)
constexpr int HE = 8;   // room in front of new symbol-stage samples: an incomplete group of 5 (<=4)

int bytes_per_sample(int fmt) { return fmt == AISGPU_FMT_CF32 ? 8 : (fmt == AISGPU_FMT_CS16 ? 4 : 2); }

} // namespace

struct aisgpu_handle {
	aisgpu_config cfg;
	int k = 0, P = 0, P96 = 0, tile = 0, bps = 8;
	// Rates the reference serves through DSP::Upsample (non-bucket rates, Model.cpp:134-149) or DSP::DownsampleKFilter
	// (288 kS/s, Model.cpp:308-313) get a pre-stage that fills a ring of whole reference blocks (d_S); the front end
	// proper then runs once per block with k = the CIC stages behind the resampler ("inner" submits).
	int pre = 0;          // 0 none, 1 = [kA x Downsample2CIC5 ->] Upsample, 2 = DownsampleKFilter / 3
	int kA = 0, PA = 0;   // pre-stage: CIC stages in front of the resampler and their warm-up history (input samples)
	int in_fmt = 0;       // sample format the front end proper reads (CF32 behind a pre-stage)
	int outer_N = 0;      // submit length of a pre-stage engine (must not change: the reference's block sizes depend on it)
	int blk = 0;          // reference block length entering the front end proper (L_us or 8192)
	int s_cap = 0;        // ring capacity in samples (whole blocks)
	long long s_stride = 0, d0_stride = 0;
	long long s_produced = 0, s_consumed = 0; // ring samples written / handed to the front end proper
	float us_alpha = 0.0f, us_inc = 1.0f;     // Upsample::alpha / increment (DSP.h:165)
	int dsk_first = 0;                         // DownsampleKFilter::idx_in
	unsigned char *d_ptail[2] = { nullptr, nullptr };
	int ptail_cur = 0;
	float2 *d_D0 = nullptr, *d_S = nullptr, *d_S2 = nullptr; // d_S2: the 96 kS/s ring behind Upsample -> DownsampleKFilter
	float2 *d_ptail2[2] = { nullptr, nullptr };
	int ptail2_cur = 0, us_blk = 0, s2_cap = 0;
	long long s2_stride = 0, s2_produced = 0, s2_consumed = 0;
	int *d_us_src = nullptr;
	float *d_us_alpha = nullptr;
	FeParams fe_pre;
	int pre_tile = 0;
	long long msg_chunk = 0; // ordinal of the caller's submit (what frames are tagged with)
	long long pre_tap0 = 0, pre_tap1 = 0, pre2_tap0 = 0, pre2_tap1 = 0; // resampler outputs of the last submit (ring positions), for AISGPU_TAP_PRE / _PRE2
	int obps = 8;            // bytes per sample of the caller's format (bps: of what the front end proper reads)
	int inner_max = 0;       // longest block the front end proper can be handed
	int use_fdc = 0;
	int fp_ds = 0; // integer CIC front end (DS_UINT16 x 4, DSP.cpp:499-665): CU8 @1536K with -go FP_DS on
	float fdc_alpha = 0, fdc_beta = 1;
	int rows = 0;
	int max_n48 = 0;
	int fe_warps = 4, fe_tile = 0, fe_ctas = 4096;
	int fe_st = 1, st_L = 0, st_ring = 0, st_kmax = 7; // AISGPU_FE_ST=0 disables the per-thread streaming kernel; AISGPU_ST_L: lanes per stream (0: the launcher plans); AISGPU_ST_NB: ring depth 3 | 5 (0: per chain)
	int cf_rows = 8; // AISGPU_CF_ROWS: rows per CTA of the fused CGF kernel (4 or 8); 8 halves the chain warp's instructions: 0.427 vs 0.433 ms per step (ModelDefault, bench.py A/B, twice)
	int dec_rpw = 6, decoder = 3; // rows per warp / which decoder kernel // front-end launch shape (tunable through AISGPU_FE_WARPS / _TILE / _CTAS)
	// fe_stream: front end + input history; stream: everything behind the 48 kHz buffers (the stream handed to callers
	// for timing).  The front end of submit c+1 overlaps the back end of submit c; Cbuf is double buffered for that.
	cudaStream_t stream = nullptr, copy_stream = nullptr, fe_stream = nullptr;
	// Back-end stages of consecutive submits are pipelined: submit c runs on be_streams[c & 1] (be_streams[0] == stream);
	// stage s of submit c waits for stage s of submit c-1 (its carried state) through ev_stage[s][(c-1) & 1]; the buffers
	// between stages are double buffered (index c & 1).  With taps enabled everything stays on one stream.
	cudaStream_t be_streams[2] = { nullptr, nullptr }, bs = nullptr;
	static const int NSTAGE = 7; // 0 estimate / fm+fir, 1 + 2 phasor chain + derot + fir (+Ec carry), 3 phase search, 4 decode, 5 carry of Cbuf, 6 Challenger FM branch (reads Ed)
	cudaEvent_t ev_stage[NSTAGE][2] = {};
	bool stage_rec[NSTAGE][2] = {};
	// Ec (FIR17 output, coherent chains) is double buffered: the fused derotation kernel of submit c + 1 writes one buffer while the
	// phase search of submit c still reads the other (with a single buffer the chain fused -> phase search of ALL submits was one
	// serial sequence).  ec_cur: the buffer the next block of symbols goes to; ev_ec_read[i]: the last reader of buffer i.
	int ec_cur = 0, ec_last = 0;
	cudaEvent_t ev_ec_read[2] = { nullptr, nullptr };
	bool ec_read_rec[2] = { false, false };
	cudaEvent_t ev_join = nullptr;
	int pb = 0; // buffer parity of the submit being enqueued
	static const int NC = 3; // ring of 48 kHz buffers: the front end may run two submits ahead of the back end
	cudaEvent_t ev_fe_done[3] = { nullptr, nullptr, nullptr }, ev_be_done[3] = { nullptr, nullptr, nullptr };
	bool be_recorded[3] = { false, false, false };
	static const int NEV = 128;
	cudaEvent_t ev_fe0s[128] = { nullptr }, ev_fe1s[128] = { nullptr };
	cudaEvent_t ev_copy[2] = { nullptr, nullptr }, ev_done[2] = { nullptr, nullptr };
	bool fe_timed = false;
	// input staging for host submits
	unsigned char *d_in[2] = { nullptr, nullptr };
	int in_cur = 0;
	bool in_used[2] = { false, false };
	unsigned char *d_tail[2] = { nullptr, nullptr };
	int tail_cur = 0;
	// Rotate
	// Rotate: table c lives in slot c % 3, rot state after chunk c in d_rot_state[1 + c % 3] (slot 0 = initial state);
	// tables are produced on side_stream, one submit ahead when the chunk length repeats
	float2 *d_rot[3] = { nullptr, nullptr, nullptr };
	int rot_n96[3] = { 0, 0, 0 };
	long long rot_ready_chunk = -1; // newest chunk whose table has been enqueued on side_stream
	int rot_cur = 0;
	float2 *d_rot_state = nullptr;  // [4]
	cudaStream_t side_stream = nullptr;
	cudaEvent_t ev_rot[3] = { nullptr, nullptr, nullptr }, ev_k1[3] = { nullptr, nullptr, nullptr };
	bool k1_recorded[3] = { false, false, false };
	float2 mult;
	// 48 kHz channel buffer
	float2 *d_C2[3] = { nullptr, nullptr, nullptr };
	int c_last = 0; // buffer the last submit's front end wrote
	long long c_stride = 0;
	int c_hist = 0; // samples kept in front of HC
	// CGF
	long long cgf_abs = 0;
	int *d_stepidx2[2] = { nullptr, nullptr };
	float2 *d_steptab = nullptr, *d_omega = nullptr, *d_cgf_rot = nullptr;
	float *d_ppmtab = nullptr;
	long long r_stride = 0;
	float2 *d_fir_hist[2] = { nullptr, nullptr };
	int fir_cur = 0;
	float2 *d_tap_cgf = nullptr;
	// symbol stage
	float2 *d_Ec2[2] = { nullptr, nullptr };
	float *d_Ef2[2] = { nullptr, nullptr };
	long long e_stride = 0;
	int e_left = 0;
	long long e_abs = 0;
	PsState *d_ps = nullptr;
	float *d_ps_mem = nullptr;
	long long *d_dbg = nullptr; // AISGPU_DEBUG=1: per-row decoder counters (tap 6)
	uint32_t *d_dbits2[2] = { nullptr, nullptr };
	float *d_lvl2[2] = { nullptr, nullptr };
	int dwords = 0;
	DecState *d_dec = nullptr;
	uint32_t *d_dec_data = nullptr;
	PllState *d_pll = nullptr;
	V2State *d_v2 = nullptr;     // V2 engine: per-row state
	float2 *d_Ed = nullptr;      // ModelChallenger: derotated 48 kHz samples [rows][HD + nE]
	long long ed_stride = 0;
	uint32_t *d_dbitsF[2] = { nullptr, nullptr };
	float *d_lvl_prev = nullptr; // [2][rows], double buffered by decode launch
	int lvlp_cur = 0;
	float2 *d_tap_coh = nullptr;
	float *d_tap_dec = nullptr, *d_tap_fm = nullptr;
	int *d_tap_cnt = nullptr;
	// frames
	// Circular frame ring.  Tickets (frames emitted since creation) only grow; the host owns `drained`.  Every submit
	// records, in stream order behind its decoders, the ring head into a pinned slot plus an event: aisgpu_poll_upto waits
	// for one submit only and later submits keep running.
	FrameRec *d_ring = nullptr;
	unsigned long long *d_ring_head = nullptr;
	int ring_cap = 0;
	unsigned long long drained = 0;           // tickets delivered to (or dropped for) the host
	static const int NT = 1024;               // submits whose completion record is kept
	unsigned long long *pin_head = nullptr;   // [NT] pinned: ring head after submit t (slot t % NT)
	cudaEvent_t ev_ticket[1024] = { nullptr };
	cudaEvent_t ev_mark = nullptr;
	std::deque<unsigned long long> launch_limit; // per undrained submit: drained-at-launch + ring_cap (what its kernels were given)
	long long polled_ticket = -1;             // newest submit whose frames have been drained
	bool overflow_pending = false;
	std::vector<FrameRec> h_ring;
	std::vector<aisgpu_msg> out_queue;
	size_t out_pos = 0;
	std::vector<int> seq; // per-stream multi-sentence sequence id (Message.cpp:28-39 is process-global in the reference)
	// last-submit geometry (for taps)
	int last_n = 0, last_n48 = 0, last_nE = 0, last_nsym = 0, last_e_begin = 0, last_launches = 0;
	uint64_t counters[8] = { 0 };
	long long chunk = 0;
	float last_fe_ms = -1.0f;
	std::string err;
	int poisoned = 0; // a CUDA failure inside a submit leaves the carried state half-advanced: every later call returns this code
	FeParams fe;
	// pinned double buffer of the Upsample (input index, alpha) schedule
	int *pin_us_src[2] = { nullptr, nullptr };
	float *pin_us_alpha[2] = { nullptr, nullptr };
	cudaEvent_t ev_us[2] = { nullptr, nullptr };
	bool us_used[2] = { false, false };
	int us_cur = 0, us_cap = 0;
	// NCCL (resolved at run time): communicator of the job's ranks, device scratch of the counter all-reduce
	void *nccl_lib = nullptr, *nccl_comm = nullptr;
	unsigned long long *d_counts = nullptr;
};

namespace {

#define CU(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess) {                                                                   \
			char b_[256];                                                                          \
			snprintf(b_, sizeof(b_), "%s:%d %.120s: %s", "aisgpu.cu", __LINE__, #call, cudaGetErrorString(e_)); \
			h->err = b_;                                                                           \
			return AISGPU_ECUDA;                                                                   \
		}                                                                                          \
	} while (0)

template <typename T>
int dalloc(aisgpu_handle *h, T **p, size_t n) {
	{
		const cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
		if (e == cudaErrorMemoryAllocation) {
			char b[160];
			snprintf(b, sizeof(b), "out of device memory allocating %zu bytes", n * sizeof(T));
			h->err = b;
			(void)cudaGetLastError();
			*p = nullptr;
			return AISGPU_ENOMEM;
		}
		CU(e);
	}
	CU(cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
	return 0;
}

// Model.cpp:129-338: which bucket, how many CIC stages, droop taps, resampler.  Returns <0 when unsupported.
int plan_frontend(aisgpu_handle *h) {
	const int sr = h->cfg.sample_rate;
	if (sr < 96000 || sr > 12288000) {
		h->err = "Model: sample rate must be between 96K and 12288K (inclusive).";
		return AISGPU_EINVAL;
	}
	static const int rates_nodsk[] = { 96000, 192000, 288000, 384000, 768000, 1536000, 3072000, 6144000, 12288000 };                           // Model.cpp:129
	static const int rates_dsk[] = { 96000, 192000, 288000, 384000, 576000, 768000, 1152000, 1536000, 2304000, 3072000, 6144000, 12288000 }; // Model.cpp:130 (-go DSK on)
	int bucket = 0;
	if (h->cfg.dsk) {
		for (int b : rates_dsk)
			if (b >= sr) { bucket = b; break; }
	}
	else {
		for (int b : rates_nodsk)
			if (b >= sr) { bucket = b; break; }
	}
	const bool interp = bucket != sr; // "sample rate ...K upsampled to ...K." (Model.cpp:146-147)
	h->pre = 0;
	h->kA = 0;
	h->fp_ds = 0;
	h->in_fmt = h->cfg.format;
	int k_total = 0;
	bool dsk_family = false;
	if (bucket == 288000 || bucket == 576000 || bucket == 1152000 || bucket == 2304000) {
		// kA x Downsample2CIC5 -> [Upsample ->] DownsampleKFilter(BlackmanHarris_28_3, 3) -> ROT, no droop filter
		// (Model.cpp:208-218, 248-258, 278-288, 308-313)
		dsk_family = true;
		for (int b = bucket; b > 288000; b >>= 1) h->kA++;
		h->pre = interp ? 3 : (h->kA ? 4 : 2); // 2: DSK on the raw input, 3: [CIC ->] Upsample -> DSK, 4: CIC -> DSK
		if (interp) h->us_inc = (float)sr / (float)bucket;
		if (h->pre != 2) {
			const int pa = 5 * ((1 << h->kA) - 1);
			const int g = std::max(4, 4 << h->kA);
			h->PA = std::max(g, (pa + g - 1) / g * g);
		}
		h->k = 0;
		h->blk = 8192; // DownsampleKFilter::outputSize (DSP.h:193)
		h->in_fmt = AISGPU_FMT_CF32;
	}
	else {
		for (int b = bucket; b > 96000; b >>= 1) k_total++;
		if (interp) { // Upsample sits in front of the last min(k, 2) CIC stages (Model.cpp:183-189 and siblings)
			const int post = k_total < 2 ? k_total : 2;
			h->pre = 1;
			h->kA = k_total - post;
			h->k = post;
			h->in_fmt = AISGPU_FMT_CF32;
			h->us_inc = (float)sr / (float)bucket; // Upsample::setParams (DSP.h:174-178)
			int pa = 5 * ((1 << h->kA) - 1);       // history a kA-stage CIC cascade needs (input samples)
			const int g = std::max(4, 4 << h->kA);    // a whole number of super-steps of the streaming kernel (and even tiles at every level)
			h->PA = std::max(g, (pa + g - 1) / g * g);
		}
		else h->k = k_total;
	}
	if (h->cfg.fp_ds) { // -go FP_DS on: only the exact 1536K bucket has an integer front end, and it is fed by convert.outCU8 (Model.cpp:222-236)
		if (sr == 1536000 && h->cfg.format == AISGPU_FMT_CU8) h->fp_ds = 1;
		else if (sr == 1536000) {
			h->err = "FP_DS on: the integer front end (Downsample16_CU8, Model.cpp:233-236) needs CU8 input";
			return AISGPU_EINVAL;
		}
	}
	(void)dsk_family;
	h->use_fdc = (h->cfg.droop && k_total > 0) ? 1 : 0;
	float a = 0.0f;
	switch (bucket) {
	case 12288000: case 6144000: a = -2.0f; break;
	case 3072000: a = -1.5f; break;
	case 1536000: case 768000: a = -1.2f; break;
	case 384000: a = -1.1f; break;
	case 192000: a = -0.8f; break;
	default: break;
	}
	h->fdc_alpha = a;
	h->fdc_beta = 1 - 2 * a; // DSP.h:293-297
	// history needed in input samples: h_0 = 17 (FDC 2 + DS2 5 + FCIC5 2*5), h_l = 2 h_{l-1} + 5
	const int k = h->k;
	int hk = 17;
	for (int i = 0; i < k; i++) hk = 2 * hk + 5;
	const int q = 1 << (k + 2);
	h->P = (hk + q - 1) / q * q;
	h->P96 = h->P >> k;
	return 0;
}

// granule of the caller's submit length: every CIC stage needs an even block (DSP.cpp:94,135)
int outer_granule(const aisgpu_handle *h) {
	if (h->fp_ds) return 16384; // 32 lane sub-segments of 512 samples: the shortest the streaming kernel takes (sub-segment >= warm-up history, 384)
	return h->pre >= 2 ? std::max(64, 1 << (h->kA + 2)) : (1 << (h->k + h->kA + 2));
}

void layout_frontend(FeParams &p, int k, int tile) {
	int off = 0;
	auto take = [&](int n) {
		int o = off;
		off += (FE_HIST + n + FE_SLACK + 1) & ~1;
		return o;
	};
	p.off_in[0] = take(tile);
	p.off_in[1] = take(tile);
	p.off_rot[0] = take(tile >> k);
	p.off_rot[1] = take(tile >> k);
	p.off_lv[0] = 0;
	for (int l = 1; l <= k; l++) p.off_lv[l] = take(tile >> l);
	p.off_up = take(tile >> k);
	p.off_dn = take(tile >> k);
	p.off_wa = take(tile >> (k + 1));
	p.off_wb = take(tile >> (k + 1));
	p.smem_f2 = off;
	p.tile = tile;
}

int launch_frontend(aisgpu_handle *h, const void *dev_in, long long stride, int N) {
	FeParams &p = h->fe;
	const int q = 1 << (h->k + 2);
	// per-CTA tile: one run of 5 outputs per thread at the first CIC stage (2 x 5 x threads input samples)
	int tile = h->fe_tile > 0 ? h->fe_tile : 320 * h->fe_warps;
	if (tile % q) tile = (tile / q + 1) * q;
	if (tile > N) tile = N;
	if (tile != p.tile) layout_frontend(p, h->k, tile);
	h->tile = tile;
	const int B = h->cfg.n_streams;
	int n_seg = (h->fe_ctas + B - 1) / B; // enough CTAs for several waves over 148 SMs
	int tiles_total = (N + tile - 1) / tile;
	if (n_seg > tiles_total) n_seg = tiles_total;
	if (n_seg < 1) n_seg = 1;
	int tiles_per_seg = (tiles_total + n_seg - 1) / n_seg;
	p.seg_len = tiles_per_seg * tile;
	n_seg = (N + p.seg_len - 1) / p.seg_len;
	p.in = dev_in;
	p.tail = h->d_tail[h->tail_cur];
	p.in_stride = stride;
	p.format = h->in_fmt;
	p.k = h->k;
	p.N = N;
	p.P = h->P;
	p.use_fdc = h->use_fdc;
	p.fdc_alpha = h->fdc_alpha;
	p.fdc_beta = h->fdc_beta;
	p.rot = h->d_rot[h->rot_cur];
	p.C = h->d_C2[h->chunk % aisgpu_handle::NC];
	p.c_stride = h->c_stride;
	p.c_off = HC;
	// 768 kS/s and above: per-thread streaming pipeline (state in registers) when the rows are 16-byte aligned; the launcher splits
	// every stream over as many lanes as make one balanced wave (st_plan, fe_stream.cuh) and declines blocks shorter than four warm-ups
	if (h->fe_st && h->k >= 3 && h->k <= h->st_kmax && ((stride * h->bps) % 16) == 0 && (((size_t)dev_in) % 16) == 0) {
		p.in = dev_in;
		p.st_B = B;
		p.st_first = h->chunk == 0 ? 1 : 0;
		// CF32 launch shape, measured live at 1024 x 131072 (gpurun probes r2b-1 .. r2b-9, per-block step times):
		//   ring of 3 chunks (104 KB per CTA) + ONE balanced wave (the planner, one CTA per SM: L = 18, 144 CTAs):
		//       ModelStandard 0.266 ms, ModelDefault 0.43 ms, Challenger 0.54 ms  <- shipped for every chain
		//   ring of 5 (174 KB) + power-of-two split L = 32 (256 CTAs, 1.73 waves; the round-2a shape):   0.294 / 0.475 / 0.56 ms
		//   ring of 5 + L = 18: 0.302-0.310 / 0.455 ms;   ring of 3 + L = 32: 0.311 ms;   ring of 3, two CTAs per SM (L = 37): 0.292-0.313 / 0.446 ms
		// The smaller ring leaves 123 KB of the SM to the back-end CTAs that run beside the front end of the next submit.
		// (Throttling the FM kernel's occupancy with unused shared memory so that it cannot crowd the front end: 0.33-0.38 ms, worse.)
		const int L = h->st_L;
		p.st_ring = h->st_ring ? h->st_ring : 3;
		p.st_cap = (h->in_fmt == 0 && !h->fp_ds) ? 1 : 0; // the integer formats run one-warp CTAs, as many per SM as fit
		const cudaError_t e = h->fp_ds ? launch_frontend_stream_fpds(p, L, h->fe_stream) : launch_frontend_stream(p, h->in_fmt, h->k, false, L, h->fe_stream);
		if (e == cudaSuccess) return 0;
		if (e != cudaErrorNotSupported) CU(e);
	}
	if (h->fp_ds) { // the integer CIC stages only exist in the streaming kernel
		h->err = "FP_DS on: n_samples must be a multiple of 16384 and the batch 16-byte aligned";
		return AISGPU_EINVAL;
	}
	const size_t smem = (size_t)p.smem_f2 * sizeof(float2);
	CU(launch_frontend_tiled(p, h->in_fmt, h->k, false, dim3(n_seg, B), smem, h->fe_stream));
	return 0;
}

// stage s of this submit may start when stage s of the previous submit (other stream) has finished
int stage_begin(aisgpu_handle *h, int s) {
	const int prev = h->pb ^ 1;
	if (h->be_streams[1] != h->be_streams[0] && h->stage_rec[s][prev]) CU(cudaStreamWaitEvent(h->bs, h->ev_stage[s][prev], 0));
	return 0;
}
int stage_end(aisgpu_handle *h, int s) {
	if (h->be_streams[1] != h->be_streams[0]) {
		CU(cudaEventRecord(h->ev_stage[s][h->pb], h->bs));
		h->stage_rec[s][h->pb] = true;
	}
	return 0;
}

int carry(aisgpu_handle *h, float2 *buf, long long stride, int src_begin, int dst_begin, int cnt) {
	if (cnt <= 0 || src_begin == dst_begin) return 0;
	CU(launch_carry_f2(buf, stride, src_begin, dst_begin, cnt, h->rows, h->bs));
	h->last_launches++;
	return 0;
}

int carry2(aisgpu_handle *h, const float2 *src, float2 *dst, long long stride, int src_begin, int dst_begin, int cnt) {
	if (cnt <= 0) return 0;
	CU(launch_carry2_f2(src, dst, stride, src_begin, dst_begin, cnt, h->rows, h->bs));
	h->last_launches++;
	return 0;
}

int run_symbols(aisgpu_handle *h, int n_new) {
	// n_new samples were appended at [HE, HE + n_new) of every row of Ec/Ef.
	// ModelDefault (ScatterPLL, DSP.h:95-117) only forwards complete groups of 5: e_left older samples sit just
	// before HE and the incomplete group at the end is carried.  ModelStandard (Deinterleave, DSP.h:65-73)
	// forwards every sample at once, so partial groups at both ends are walked with a per-phase validity test.
	if (h->cfg.model == AISGPU_MODEL_STANDARD) {
		const long long a0 = h->e_abs, a1 = a0 + n_new;
		const long long g0 = a0 - a0 % 5;
		const int nslots = (int)((a1 - g0 + 4) / 5);
		h->last_nsym = nslots;
		K3Params p;
		memset(&p, 0, sizeof(p));
		p.rows = h->rows;
		p.nsym = nslots;
		p.e_stride = h->e_stride;
		p.e_begin = HE - (int)(a0 - g0);
		p.abs_begin = g0;
		p.abs_lo = a0;
		p.abs_hi = a1;
		p.Ef = h->d_Ef2[0];
		p.dec = h->d_dec;
		p.dec_data = h->d_dec_data;
		p.ring = h->d_ring;
		p.ring_head = h->d_ring_head;
		p.ring_limit = h->drained + (unsigned long long)h->ring_cap;
		p.ring_cap = h->ring_cap;
		p.chunk = (int)h->msg_chunk;
		p.blk = (int)h->chunk;
		p.mode_level = (h->cfg.tag_mode & 1) ? 1 : 0;
		p.tap_dec = nullptr; // the decoder input samples are recorded by the FM/FIR kernel
		p.dbg = h->d_dbg;
		p.dbits = h->d_dbits2[h->pb];
		p.dwords = h->dwords;
		if (int rc = stage_begin(h, 4)) return rc;
		CU(launch_decode(0, h->decoder, h->dec_rpw, p, h->bs));
		if (int rc = stage_end(h, 4)) return rc;
		h->last_launches++;
		h->e_abs = a1;
		return 0;
	}
	const int total = h->e_left + n_new;
	const int nsym = total / 5;
	const int e_begin = HE - h->e_left;
	h->last_nsym = nsym;
	h->last_e_begin = e_begin;
	if (nsym > 0) {
		K3Params p;
		memset(&p, 0, sizeof(p));
		p.ps_ema = h->cfg.ps_ema;
		p.ps_rot0 = (int)((h->e_abs / 5) & 3); // e_abs counts from 0 at creation: symbols delivered so far
		p.rows = h->rows;
		p.nsym = nsym;
		p.e_stride = h->e_stride;
		p.e_begin = e_begin;
		p.abs_begin = h->e_abs;
		p.abs_lo = h->e_abs;
		p.abs_hi = h->e_abs + (long long)nsym * 5;
		p.Ec = h->d_Ec2[h->ec_cur];
		p.Ef = h->d_Ef2[0];
		p.ps = h->d_ps;
		p.ps_mem = h->d_ps_mem;
		p.dec = h->d_dec;
		p.dec_data = h->d_dec_data;
		p.ring = h->d_ring;
		p.ring_head = h->d_ring_head;
		p.ring_limit = h->drained + (unsigned long long)h->ring_cap;
		p.ring_cap = h->ring_cap;
		p.chunk = (int)h->msg_chunk;
		p.blk = (int)h->chunk;
		p.mode_level = (h->cfg.tag_mode & 1) ? 1 : 0;
		if (h->cfg.model == AISGPU_MODEL_DEFAULT || h->cfg.model == AISGPU_MODEL_CHALLENGER) {
			p.stepidx = h->d_stepidx2[h->pb];
			p.ppmtab = h->d_ppmtab;
			p.blk_abs0 = h->cgf_abs;
			p.nblk = n_new / CGF_N;
		}
		p.tap_dec = h->cfg.enable_taps ? h->d_tap_dec : nullptr;
		p.dbg = h->d_dbg;
		p.dbits = h->d_dbits2[h->pb];
		p.dwords = h->dwords;
		p.lvl = h->d_lvl2[h->pb];
		p.lvl_stride = h->dwords * K3_TS;
		if (int rc = stage_begin(h, 3)) return rc;
		CU(launch_phase_search(p, h->bs));
		// the incomplete group of 5 at the end moves to the front of the OTHER Ec buffer, where the next block of symbols lands
		const int nl = total - nsym * 5;
		if (carry2(h, h->d_Ec2[h->ec_cur], h->d_Ec2[h->ec_cur ^ 1], h->e_stride, e_begin + nsym * 5, HE - nl, nl)) return AISGPU_ECUDA;
		if (int rc = stage_end(h, 3)) return rc;
		CU(cudaEventRecord(h->ev_ec_read[h->ec_cur], h->bs));
		h->ec_read_rec[h->ec_cur] = true;
		if (h->cfg.model == AISGPU_MODEL_CHALLENGER) {
			// FM branch on the derotated samples: Demod::FM -> Filter 37 -> Deinterleave (Model.cpp:637-639), all new samples at once
			const long long a0 = h->e_abs + h->e_left, a1 = a0 + n_new; // absolute indices of the new samples
			Fm5Params f;
			memset(&f, 0, sizeof(f));
			f.Cbuf = h->d_Ed;
			f.c_stride = h->ed_stride;
			f.c_new = HD;
			f.n = n_new;
			f.r0 = h->e_left;
			f.nslots = (int)((a1 - h->e_abs + 4) / 5);
			f.Fbuf = h->cfg.enable_taps ? h->d_Ef2[0] : nullptr; // k_decode10 takes the decision bits
			f.f_stride = h->e_stride;
			f.f_off = HE;
			f.dbits = h->d_dbitsF[h->pb];
			f.dwords = h->dwords;
			if (int rc = stage_begin(h, 6)) return rc;
			CU(launch_fm_fir5(f, h->rows, h->bs));
			if (int rc = carry(h, h->d_Ed, h->ed_stride, HD + n_new - HD, 0, HD)) return rc; // the last HD derotated samples stay in front
			if (int rc = stage_end(h, 6)) return rc;
			p.dbits2 = h->d_dbitsF[h->pb];
			p.nslots_fm = f.nslots;
			p.lvl_prev = h->d_lvl_prev + (size_t)h->lvlp_cur * h->rows;
			p.lvl_prev_out = h->d_lvl_prev + (size_t)(h->lvlp_cur ^ 1) * h->rows;
			h->lvlp_cur ^= 1;
			p.abs_lo = a0;
			p.abs_hi = a1;
			if (int rc = stage_begin(h, 4)) return rc;
			CU(launch_decode10(h->dec_rpw == 1 ? 1 : 3, p, h->bs));
			if (int rc = stage_end(h, 4)) return rc;
			h->last_launches += 3;
		}
		else {
			if (int rc = stage_begin(h, 4)) return rc;
			CU(launch_decode(2, h->decoder, h->dec_rpw, p, h->bs));
			if (int rc = stage_end(h, 4)) return rc;
			h->last_launches += 2;
		}
	}
	const int new_left = total - nsym * 5;
	if (nsym == 0) {
		if (int rc = stage_begin(h, 3)) return rc;
		if (carry2(h, h->d_Ec2[h->ec_cur], h->d_Ec2[h->ec_cur ^ 1], h->e_stride, e_begin, HE - new_left, new_left)) return AISGPU_ECUDA;
		if (int rc = stage_end(h, 3)) return rc;
		CU(cudaEventRecord(h->ev_ec_read[h->ec_cur], h->bs));
		h->ec_read_rec[h->ec_cur] = true;
	}
	h->ec_last = h->ec_cur;
	h->ec_cur ^= 1;
	h->e_left = new_left;
	h->e_abs += (long long)nsym * 5;
	return 0;
}

// Enqueue the Rotate phasor table of chunk c (n96 samples at 96 kHz) on the side stream.
int enqueue_rot_table(aisgpu_handle *h, long long c, int n96) {
	const int slot = (int)(c % 3), prev = (int)((c + 2) % 3);
	// slot was last read by the front end of chunk c-3
	if (h->k1_recorded[slot]) CU(cudaStreamWaitEvent(h->side_stream, h->ev_k1[slot], 0));
	const float2 *prev_tail = c > 0 ? h->d_rot[prev] + h->rot_n96[prev] : nullptr;
	const float2 *state_in = c > 0 ? h->d_rot_state + 1 + prev : h->d_rot_state;
	CU(launch_rot_table(h->d_rot[slot], prev_tail, state_in, h->d_rot_state + 1 + slot, h->mult, h->P96, n96, h->side_stream));
	CU(cudaEventRecord(h->ev_rot[slot], h->side_stream));
	h->rot_n96[slot] = n96;
	h->rot_ready_chunk = c;
	return 0;
}

// One Receive() of the front end proper: N samples per stream (a whole reference block) -> frames.
int submit_common(aisgpu_handle *h, const void *dev_in, long long stride, int N) {
	const int q = 1 << (h->k + 2);
	if (N <= 0 || N > h->inner_max || (N % q) != 0) {
		char b[160];
		snprintf(b, sizeof(b), "internal: block of %d samples must be a positive multiple of %d and <= %d", N, q, h->inner_max);
		h->err = b;
		return AISGPU_ECUDA; // cannot come from the caller's arguments (check_outer has passed): poisons the handle
	}
	const int k = h->k, B = h->cfg.n_streams;
	const int n96 = N >> k, n48 = n96 >> 1;
	// ---- K0: Rotate phasor table (side stream; normally already enqueued by the previous submit) ----
	{
		const long long c = h->chunk;
		const int slot = (int)(c % 3);
		if (!(h->rot_ready_chunk == c && h->rot_n96[slot] == n96)) {
			if (int rc = enqueue_rot_table(h, c, n96)) return rc;
		}
		CU(cudaStreamWaitEvent(h->fe_stream, h->ev_rot[slot], 0));
		h->rot_cur = slot;
	}
	const int cb = (int)(h->chunk % aisgpu_handle::NC);
	float2 *Ccur = h->d_C2[cb], *Cnext = h->d_C2[(cb + 1) % aisgpu_handle::NC];
	h->c_last = cb;
	// ---- K1: fused front end (its own stream: overlaps the back end of the previous submit) ----
	if (h->be_recorded[cb]) CU(cudaStreamWaitEvent(h->fe_stream, h->ev_be_done[cb], 0)); // back end of submit c-3 still reads Cbuf[cb]
	const int evi = (int)(h->chunk % aisgpu_handle::NEV);
	CU(cudaEventRecord(h->ev_fe0s[evi], h->fe_stream));
	if (int rc = launch_frontend(h, dev_in, stride, N)) return rc;
	CU(cudaEventRecord(h->ev_fe1s[evi], h->fe_stream));
	CU(cudaEventRecord(h->ev_k1[h->chunk % 3], h->fe_stream));
	h->k1_recorded[h->chunk % 3] = true;
	h->fe_timed = true;
	h->last_launches += 2; // front end + this submit's phasor table
	// speculate that the next submit has the same length: build its phasor table now, off the critical path
	if (int rc = enqueue_rot_table(h, h->chunk + 1, n96)) return rc;
	// ---- front-end history for the next submit ----
	{
		const int nxt = h->tail_cur ^ 1;
		const int p_w = h->P * h->bps / 8;
		CU(launch_tail_update(h->d_tail[nxt], h->d_tail[h->tail_cur], dev_in, stride * h->bps / 8, (long long)N * h->bps / 8, p_w, B, h->fe_stream));
		h->tail_cur = nxt;
		h->last_launches++;
	}
	CU(cudaEventRecord(h->ev_fe_done[cb], h->fe_stream));
	h->pb = (int)(h->chunk & 1);
	h->bs = h->be_streams[h->pb];
	CU(cudaStreamWaitEvent(h->bs, h->ev_fe_done[cb], 0));
	h->last_n = N;
	h->last_n48 = n48;
	h->last_nE = 0;
	h->last_nsym = 0;
	// ---- back end (stage-pipelined over consecutive submits, see aisgpu_handle::be_streams) ----
	if (h->cfg.model == AISGPU_MODEL_V2) {
		// Engine::Receive (V2Engine.cpp:379-395): a block is decoded once the NEXT block is complete (it is the estimator's
		// lookahead), so one whole block plus the partial one wait in front of the new samples
		const int cnt = h->c_hist;
		const int total = cnt + n48;
		const int nproc = std::max(0, total / V2_BLK - 1);
		const int c_begin = HC - cnt;
		const int newcnt = total - nproc * V2_BLK;
		if (int rc = carry2(h, Ccur, Cnext, h->c_stride, c_begin + nproc * V2_BLK, HC - newcnt, newcnt)) return rc;
		h->c_hist = newcnt;
		if (nproc > 0) {
			CU(launch_v2_engine(Ccur, h->c_stride, c_begin, nproc, h->rows, h->d_v2, h->d_dec, h->d_dec_data, h->d_ring, h->d_ring_head,
								h->drained + (unsigned long long)h->ring_cap, h->ring_cap, (int)h->msg_chunk, (int)h->chunk, (h->cfg.tag_mode & 1) ? 1 : 0, h->d_omega,
								h->cfg.dd_train, h->cfg.dd_weight, h->cfg.enable_taps ? h->d_tap_cgf : nullptr, h->cfg.enable_taps ? h->d_tap_coh : nullptr,
								h->cfg.enable_taps ? h->d_tap_fm : nullptr, h->r_stride, h->bs));
			h->last_launches++;
		}
		h->last_nE = nproc * V2_BLK;
		CU(cudaEventRecord(h->ev_be_done[cb], h->bs));
		h->be_recorded[cb] = true;
	}
	else if (h->cfg.model == AISGPU_MODEL_DEFAULT || h->cfg.model == AISGPU_MODEL_CHALLENGER) {
		const int cnt = h->c_hist; // unconsumed samples in front of HC
		const int total = cnt + n48;
		const int nblk = total / CGF_N;
		const int c_begin = HC - cnt;
		const int newcnt = total - nblk * CGF_N;
		// samples that do not fill a 512-block go to the front of the next submit's buffer; the estimator of the next
		// submit only has to wait for this copy (and this one for the copy of the previous submit)
		if (int rc = stage_begin(h, 5)) return rc;
		if (int rc = carry2(h, Ccur, Cnext, h->c_stride, c_begin + nblk * CGF_N, HC - newcnt, newcnt)) return rc;
		if (int rc = stage_end(h, 5)) return rc;
		h->c_hist = newcnt;
		if (nblk > 0) {
			const int total_blocks = h->rows * nblk;
			int *stepidx = h->d_stepidx2[h->pb];
			// stepidx / dbits / lvl are double buffered by submit parity == stream, so stream order protects them
			CU(launch_cgf_estimate(Ccur, h->c_stride, c_begin, nblk, total_blocks, h->d_omega, h->cfg.afc_wide, stepidx, h->bs));
			const int nE = nblk * CGF_N;
			{ // phasor chain + derotation + FIR17 in one kernel: waits for what carries its state (stages 1, 2) and for the last reader of
				// the Ec buffer it writes (the phase search two blocks of symbols ago)
				if (int rc = stage_begin(h, 1)) return rc;
				if (int rc = stage_begin(h, 2)) return rc;
				if (h->ec_read_rec[h->ec_cur]) CU(cudaStreamWaitEvent(h->bs, h->ev_ec_read[h->ec_cur], 0));
				if (h->d_Ed) // ModelChallenger: Ed (single buffered) is free once the previous submit's FM branch has read it
					if (int rc = stage_begin(h, 6)) return rc;
				CU(launch_cgf_fused(Ccur, h->c_stride, c_begin, stepidx, h->d_steptab, h->d_cgf_rot, nblk, h->rows, h->d_fir_hist[h->fir_cur],
									h->d_fir_hist[h->fir_cur ^ 1], h->d_Ec2[h->ec_cur], h->e_stride, HE,
									h->d_Ed ? h->d_Ed + HD : (h->cfg.enable_taps ? h->d_tap_cgf : nullptr), h->d_Ed ? h->ed_stride : h->r_stride, h->cf_rows, h->bs));
				if (int rc = stage_end(h, 1)) return rc;
				if (int rc = stage_end(h, 2)) return rc;
			}
			h->fir_cur ^= 1;
			h->last_launches += 2;
			h->last_nE = nE;
		}
		CU(cudaEventRecord(h->ev_be_done[cb], h->bs)); // last reader of Cbuf[cb]
		h->be_recorded[cb] = true;
		if (nblk > 0) {
			if (int rc = run_symbols(h, nblk * CGF_N)) return rc;
			h->cgf_abs += nblk * CGF_N;
		}
	}
	else {
		if (int rc = stage_begin(h, 5)) return rc;
		if (int rc = carry2(h, Ccur, Cnext, h->c_stride, HC + n48 - FIRF_T, HC - FIRF_T, FIRF_T)) return rc; // FM + FIR history
		if (int rc = stage_end(h, 5)) return rc;
		{
			// the 5-phase deinterleaver's slots are aligned to absolute sample indices (DSP.h:65-73)
			const long long a0 = h->e_abs, a1 = a0 + n48;
			const long long g0 = a0 - a0 % 5;
			const int nslots = (int)((a1 - g0 + 4) / 5);
			Fm5Params f;
			memset(&f, 0, sizeof(f));
			f.Cbuf = Ccur;
			f.c_stride = h->c_stride;
			f.c_new = HC;
			f.n = n48;
			f.r0 = h->cfg.model == AISGPU_MODEL_STANDARD ? (int)(a0 - g0) : 0;
			f.nslots = nslots;
			// the filtered samples themselves are only read by k_base, the bit-serial cross-check decoder and the taps
			f.Fbuf = (h->cfg.model == AISGPU_MODEL_BASE || h->decoder == 1 || h->cfg.enable_taps) ? h->d_Ef2[0] : nullptr;
			f.f_stride = h->e_stride;
			f.f_off = HE;
			f.dbits = h->d_dbits2[h->pb];
			f.dwords = h->dwords;
			f.tap_fm = h->cfg.enable_taps ? h->d_tap_fm : nullptr;
			f.tap_stride = h->r_stride;
			f.tap_dec = (h->cfg.enable_taps && h->cfg.model == AISGPU_MODEL_STANDARD) ? h->d_tap_dec : nullptr;
			if (int rc = stage_begin(h, 0)) return rc; // Ef (single buffered) is only read by taps / k_base, which do not pipeline
			CU(launch_fm_fir5(f, h->rows, h->bs));
			if (int rc = stage_end(h, 0)) return rc;
		}
		CU(cudaEventRecord(h->ev_be_done[cb], h->bs));
		h->be_recorded[cb] = true;
		h->last_launches++;
		h->last_nE = n48;
		if (h->cfg.model == AISGPU_MODEL_STANDARD) {
			if (int rc = run_symbols(h, n48)) return rc;
		}
		else {
			CU(launch_base(h->d_Ef2[0], h->e_stride, HE, n48, h->rows, h->d_pll, h->d_dec, h->d_dec_data, h->d_ring, h->d_ring_head, h->drained + (unsigned long long)h->ring_cap, h->ring_cap, (int)h->msg_chunk,
						   (int)h->chunk, h->cfg.enable_taps ? h->d_tap_dec : nullptr, h->cfg.enable_taps ? h->d_tap_cnt : nullptr, h->bs));
			h->last_launches++;
		}
	}
	h->chunk++;
	return 0;
}

// DownsampleKFilter over N input samples per stream (any format) -> ring of 96 kS/s samples
int run_dsk(aisgpu_handle *h, const void *in, long long stride, int fmt, int N, const void *tail, float2 *S, long long s_stride, long long &produced, int cap) {
	const int B = h->cfg.n_streams;
	const int first = h->dsk_first;
	const int n_out = first < N ? (N - first + 2) / 3 : 0;
	if (n_out > 0) {
		CU(launch_dsk(fmt, in, stride, tail, 32, first, n_out, B, S, s_stride, produced, cap, h->fe_stream));
		h->last_launches++;
	}
	h->dsk_first = first + 3 * n_out - N;
	produced += n_out;
	return 0;
}

int sync_backend(aisgpu_handle *h) {
	CU(cudaStreamSynchronize(h->be_streams[0]));
	if (h->be_streams[1] != h->be_streams[0]) CU(cudaStreamSynchronize(h->be_streams[1]));
	return 0;
}

int check_outer(aisgpu_handle *h, int N) {
	const int q = outer_granule(h);
	if (N <= 0 || N > h->cfg.max_chunk_samples || (N % q) != 0) {
		char b[160];
		snprintf(b, sizeof(b), "n_samples=%d must be a positive multiple of %d and <= max_chunk_samples=%d", N, q, h->cfg.max_chunk_samples);
		h->err = b;
		return AISGPU_EINVAL;
	}
	if ((h->pre == 1 || h->pre == 3) && h->outer_N && N != h->outer_N) {
		h->err = "at an interpolated sample rate every submit must have the same length (DSP::Upsample re-blocks by it, DSP.cpp:203)";
		return AISGPU_EINVAL;
	}
	return 0;
}

int mark_ticket(aisgpu_handle *h, long long t);

// The caller's Receive(): N samples per stream in the caller's format.
int submit_outer(aisgpu_handle *h, const void *dev_in, long long stride, int N) {
	if (int rc = check_outer(h, N)) return rc;
	h->last_launches = 0;
	h->msg_chunk = (long long)h->counters[3];
	const int B = h->cfg.n_streams;
	int rc = 0;
	h->pre_tap0 = h->s_produced;
	h->pre2_tap0 = h->s2_produced;
	if (h->pre == 0) rc = submit_common(h, dev_in, stride, N);
	else {
		const int cur = h->ptail_cur, nxt = cur ^ 1;
		int tail_len = 0;
		if (h->pre == 1 || h->pre == 3 || h->pre == 4) { // kA x Downsample2CIC5 -> Upsample | DownsampleKFilter (Model.cpp:183-189, 208-218; DSP.cpp:192-212)
			const int L = N >> h->kA;
			if (!h->outer_N && h->pre != 4) {
				h->outer_N = N;
				h->us_blk = L;
				if (h->pre == 1) h->blk = L;
				h->s_cap = 4 * L;
			}
			FeParams &pp = h->fe_pre;
			int tile = 1280;
			if (tile > N) tile = N;
			if (tile != pp.tile) layout_frontend(pp, h->kA, tile);
			int n_seg = (h->fe_ctas + B - 1) / B;
			const int tiles_total = (N + tile - 1) / tile;
			n_seg = std::max(1, std::min(n_seg, tiles_total));
			pp.seg_len = (tiles_total + n_seg - 1) / n_seg * tile;
			n_seg = (N + pp.seg_len - 1) / pp.seg_len;
			pp.in = dev_in;
			pp.tail = h->d_ptail[cur];
			pp.in_stride = stride;
			pp.format = h->cfg.format;
			pp.k = h->kA;
			pp.N = N;
			pp.P = h->PA;
			pp.use_fdc = 0;
			pp.rot = nullptr;
			pp.C = nullptr;
			pp.D0 = h->d_D0;
			pp.d0_stride = h->d0_stride;
			pp.d0_off = 2;
			const size_t smem = (size_t)pp.smem_f2 * sizeof(float2);
			dim3 grid(n_seg, B);
			bool st_done = false;
			if (h->fe_st && h->kA >= 3 && h->kA <= 5 && ((stride * h->obps) % 16) == 0 && (((size_t)dev_in) % 16) == 0) {
				pp.st_B = B;
				// the decimation in front of the resampler runs the same launch shape as the front end proper (four-warp CTAs, ring of 3, one
				// balanced wave): configs[2] (4096 x 65536 @6 MSPS) 0.70 -> 0.66 ms per step, 1024 x 393216 @6 MSPS 0.91 -> 0.83 ms against
				// one-warp CTAs with 16-sample chunks
				pp.st_ring = h->st_ring ? h->st_ring : 3;
				pp.st_cap = h->cfg.format == 0 ? 1 : 0;
				const cudaError_t e = launch_frontend_stream(pp, h->cfg.format, h->kA, true, h->st_L, h->fe_stream);
				if (e == cudaSuccess) st_done = true;
				else if (e != cudaErrorNotSupported) CU(e);
			}
			if (!st_done) CU(launch_frontend_tiled(pp, h->cfg.format, h->kA, true, grid, smem, h->fe_stream));
			if (rc) return rc;
			tail_len = h->PA;
			if (h->pre == 4) { // the level-kA stream goes straight through DownsampleKFilter into the 96 kS/s ring
				const int c2 = h->ptail2_cur;
				if ((rc = run_dsk(h, h->d_D0 + 2, h->d0_stride, AISGPU_FMT_CF32, L, h->d_ptail2[c2], h->d_S, h->s_stride, h->s_produced, h->s_cap))) return rc;
				CU(launch_tail_update(h->d_ptail2[c2 ^ 1], h->d_ptail2[c2], h->d_D0 + 2, h->d0_stride, (long long)L, 32, B, h->fe_stream));
				h->ptail2_cur = c2 ^ 1;
				h->last_launches += 2;
			}
			else {
			// replay Upsample's float accumulator: one (input index, alpha) pair per output (DSP.cpp:196-209).  The table goes
			// through a pinned double buffer, so the copy is a true asynchronous one and the caller's thread never waits
			// for the front-end stream here.
			const int ub = h->us_cur;
			if (h->us_used[ub]) CU(cudaEventSynchronize(h->ev_us[ub])); // the copy issued two submits ago has read this buffer
			int *us_src = h->pin_us_src[ub];
			float *us_al = h->pin_us_alpha[ub];
			float alpha = h->us_alpha;
			const float inc = h->us_inc;
			int M = 0;
			for (int i = 0; i < L; i++) {
				do {
					if (M >= h->us_cap) {
						h->err = "Upsample schedule overflow";
						return AISGPU_ECUDA;
					}
					us_src[M] = i;
					us_al[M++] = alpha;
					alpha += inc;
				} while (alpha < 1.0f);
				alpha -= 1.0f;
			}
			h->us_alpha = alpha;
			CU(cudaMemcpyAsync(h->d_us_src, us_src, (size_t)M * sizeof(int), cudaMemcpyHostToDevice, h->fe_stream));
			CU(cudaMemcpyAsync(h->d_us_alpha, us_al, (size_t)M * sizeof(float), cudaMemcpyHostToDevice, h->fe_stream));
			CU(cudaEventRecord(h->ev_us[ub], h->fe_stream));
			h->us_used[ub] = true;
			h->us_cur ^= 1;
			CU(launch_upsample(h->d_D0, h->d0_stride, 2, h->d_us_src, h->d_us_alpha, M, B, h->d_S, h->s_stride, h->s_produced, h->s_cap, h->fe_stream));
			CU(launch_d0_carry(h->d_D0, h->d0_stride, 2, L, B, h->fe_stream));
			h->s_produced += M;
			h->last_launches += 3;
			}
		}
		else { // DownsampleKFilter(BlackmanHarris_28_3, 3) (Model.cpp:308-313; DSP.cpp:160-189)
			tail_len = 32;
			if ((rc = run_dsk(h, dev_in, stride, h->cfg.format, N, h->d_ptail[cur], h->d_S, h->s_stride, h->s_produced, h->s_cap))) return rc;
		}
		{ // raw-format history of the pre-stage for the next submit
			const int p_w = tail_len * h->obps / 8;
			CU(launch_tail_update(h->d_ptail[nxt], h->d_ptail[cur], dev_in, stride * h->obps / 8, (long long)N * h->obps / 8, p_w, B, h->fe_stream));
			h->ptail_cur = nxt;
			h->last_launches++;
		}
		if (h->pre == 3) { // every complete Upsample block goes through DownsampleKFilter into the 96 kS/s ring
			while (h->s_produced - h->s_consumed >= h->us_blk) {
				const int slot = (int)(h->s_consumed % h->s_cap);
				const int c2 = h->ptail2_cur;
				if ((rc = run_dsk(h, h->d_S + slot, h->s_stride, AISGPU_FMT_CF32, h->us_blk, h->d_ptail2[c2], h->d_S2, h->s2_stride, h->s2_produced, h->s2_cap))) return rc;
				CU(launch_tail_update(h->d_ptail2[c2 ^ 1], h->d_ptail2[c2], h->d_S + slot, h->s_stride, (long long)h->us_blk, 32, B, h->fe_stream));
				h->ptail2_cur = c2 ^ 1;
				h->s_consumed += h->us_blk;
			}
			while (h->s2_produced - h->s2_consumed >= h->blk) {
				const int slot = (int)(h->s2_consumed % h->s2_cap);
				if ((rc = submit_common(h, h->d_S2 + slot, h->s2_stride, h->blk))) return rc;
				h->s2_consumed += h->blk;
			}
		}
		else {
			// hand every complete reference block to the front end proper
			while (h->s_produced - h->s_consumed >= h->blk) {
				const int slot = (int)(h->s_consumed % h->s_cap);
				if ((rc = submit_common(h, h->d_S + slot, h->s_stride, h->blk))) return rc;
				h->s_consumed += h->blk;
			}
		}
	}
	if (rc) return rc;
	h->pre_tap1 = h->s_produced;
	h->pre2_tap1 = h->s2_produced;
	if (int rc2 = mark_ticket(h, (long long)h->counters[3])) return rc2;
	h->counters[2] += (uint64_t)N;
	h->counters[3] += 1;
	return 0;
}

// The payload of a frame as NMEA six-bit letters, one pass over the bit stream (what Message::getLetter does letter by
// letter, Message.cpp:643-662): letter i = bits [6i, 6i + 6), MSB first; bits past the message end read as 0; a letter that
// would cross bit 1064 is the NUL byte the reference returns there; value v prints as v + 48 (v < 40) or v + 56.
int armour_payload(const uint8_t *data, int nbits, char *out) {
	const int n = (nbits + 5) / 6;
	uint32_t window = 0; // the next `have` unread bits, right-aligned
	int have = 0, next_byte = 0;
	for (int i = 0; i < n; i++) {
		if (have < 6) {
			window = (window << 8) | data[next_byte++];
			have += 8;
		}
		have -= 6;
		unsigned v = (window >> have) & 0x3Fu;
		const int past = 6 * (i + 1) - nbits; // bits of this letter beyond the end of the message
		if (past > 0) v &= 0x3Fu << past;
		out[i] = 6 * (i + 1) > 1064 ? (char)0 : (char)(v + (v < 40 ? 48 : 56));
	}
	return n;
}

bool msg_validate(const uint8_t *d, int length) { // Message.cpp:398-413
	static const int ml[28] = { 149, 149, 149, 168, 418, 88, 72, 56, 168, 70, 168, 72, 40, 40, 88, 92, 80, 168, 312, 70, 271, 145, 154, 160, 72, 60, 96, 168 };
	if (length == 0) return true;
	if (length > 1064) return false;
	unsigned t = d[0] >> 2;
	if (t < 1 || t > 28) return false;
	return length >= ml[t - 1];
}

// Message::buildNMEA (Message.cpp:569-631): "!AIVDM,<sentences>,<index>,<seq>,<channel>,<up to 56 letters>,<fill>*<checksum>";
// own-ship frames read !AIVDO; the sequence id (Message::nextSeqId, Message.cpp:28-39; one counter per stream here instead of
// one per process) only exists for multi-sentence messages; fill bits are reported on the last sentence.
void build_nmea(aisgpu_msg &m, int own_mmsi, int *seq_counter) {
	char letters[180];
	const int nletters = armour_payload(m.data, m.nbits, letters);
	const int nsent = nletters ? (nletters + 55) / 56 : 1;
	const uint32_t mmsi = ((uint32_t)m.data[1] << 22) | ((uint32_t)m.data[2] << 14) | ((uint32_t)m.data[3] << 6) | (m.data[4] >> 2);
	char seq = 0;
	if (nsent > 1) {
		seq = (char)('0' + *seq_counter);
		*seq_counter = (*seq_counter + 1) % 10;
	}
	m.n_sentences = nsent;
	for (int s = 0; s < nsent && s < 4; s++) {
		char *line = m.nmea[s];
		int at = 0;
		uint8_t sum = 0; // XOR of everything between '!' and '*'
		auto put = [&](char ch) { line[at++] = ch; sum ^= (uint8_t)ch; };
		line[at++] = '!';
		for (const char *t = (own_mmsi == (int)mmsi) ? "AIVDO," : "AIVDM,"; *t; t++) put(*t);
		put((char)('0' + nsent));
		put(',');
		put((char)('1' + s));
		put(',');
		if (seq) put(seq);
		put(',');
		if (m.channel != '?') put(m.channel);
		put(',');
		const int first = 56 * s, count = std::min(56, nletters - first);
		for (int k = 0; k < count; k++) put(letters[first + k]);
		put(',');
		put((char)('0' + (s == nsent - 1 ? 6 * nletters - m.nbits : 0)));
		line[at++] = '*';
		line[at++] = "0123456789ABCDEF"[sum >> 4];
		line[at++] = "0123456789ABCDEF"[sum & 15];
		line[at] = 0;
		m.nmea_len[s] = at; // a 1064-bit message ends in a NUL letter, so strlen() is not enough
	}
}

// Copies the frames with tickets [a, b) to the end of h_ring; tickets at or above `limit` were dropped by the kernels.
int fetch_frames(aisgpu_handle *h, unsigned long long a, unsigned long long b, unsigned long long limit) {
	const unsigned long long hi = std::min(b, std::max(a, limit));
	if (b > hi) {
		h->counters[4] += (uint64_t)(b - hi);
		h->overflow_pending = true;
	}
	const unsigned long long cap = (unsigned long long)h->ring_cap;
	while (a < hi) {
		const size_t off = (size_t)(a % cap);
		const size_t n = (size_t)std::min<unsigned long long>(hi - a, cap - off);
		const size_t at = h->h_ring.size();
		h->h_ring.resize(at + n);
		CU(cudaMemcpy(h->h_ring.data() + at, h->d_ring + off, n * sizeof(FrameRec), cudaMemcpyDeviceToHost));
		a += n;
	}
	return 0;
}

// Records, behind everything enqueued for submit `t`, the ring head (into a pinned slot) and the completion event.
int mark_ticket(aisgpu_handle *h, long long t) {
	cudaStream_t st = h->bs ? h->bs : h->stream;
	CU(cudaEventRecord(h->ev_mark, h->fe_stream)); // a pre-stage submit may have launched nothing behind the front-end stream
	CU(cudaStreamWaitEvent(st, h->ev_mark, 0));
	const int slot = (int)(t % aisgpu_handle::NT);
	// With two back-end streams the snapshots of consecutive submits sit on different streams: chain them, so that the recorded
	// heads are monotonic in t and record t also covers everything submit t-1 wrote (a submit without back-end work would otherwise
	// take its snapshot while the previous submit's decoders are still running)
	if (t > 0 && h->be_streams[1] != h->be_streams[0]) CU(cudaStreamWaitEvent(st, h->ev_ticket[(t - 1) % aisgpu_handle::NT], 0));
	CU(cudaMemcpyAsync(&h->pin_head[slot], h->d_ring_head, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CU(cudaEventRecord(h->ev_ticket[slot], st));
	h->launch_limit.push_back(h->drained + (unsigned long long)h->ring_cap);
	return 0;
}

// Frames of the submits (polled_ticket, upto] -> out_queue.  upto < 0: everything, after a full synchronisation.
int drain_ring(aisgpu_handle *h, long long upto) {
	const long long latest = (long long)h->counters[3] - 1;
	if (latest < 0) return 0;
	if (upto < 0 || upto > latest) {
		CU(cudaStreamSynchronize(h->fe_stream));
		if (int rc = sync_backend(h)) return rc;
		upto = latest;
	}
	if (upto <= h->polled_ticket) return 0;
	if (latest - upto >= aisgpu_handle::NT) upto = latest; // its completion record has been recycled: wait for the newest one
	CU(cudaEventSynchronize(h->ev_ticket[upto % aisgpu_handle::NT]));
	h->h_ring.clear();
	const long long first = h->polled_ticket + 1;
	if (latest - first < aisgpu_handle::NT) { // every submit's own record is still there: exact limits per submit
		for (long long t = first; t <= upto; t++) {
			const unsigned long long head = std::max(h->drained, (unsigned long long)h->pin_head[t % aisgpu_handle::NT]);
			if (int rc = fetch_frames(h, h->drained, head, h->launch_limit.front())) return rc;
			h->drained = head;
			h->launch_limit.pop_front();
		}
	}
	else { // more submits than records since the last poll: they all ran against the oldest one's limit or a later (larger) one
		const unsigned long long head = std::max(h->drained, (unsigned long long)h->pin_head[upto % aisgpu_handle::NT]);
		if (int rc = fetch_frames(h, h->drained, head, h->launch_limit.front())) return rc;
		h->drained = head;
		for (long long t = first; t <= upto; t++) h->launch_limit.pop_front();
	}
	h->polled_ticket = upto;
	if (h->h_ring.empty()) return 0;
	// reference emission order: per submit, stream-major, channel A (ROT.up) before B (DSP.cpp:312-313), then time
	std::stable_sort(h->h_ring.begin(), h->h_ring.end(), [](const FrameRec &a, const FrameRec &b) {
		if (a.chunk != b.chunk) return a.chunk < b.chunk;
		if (a.blk != b.blk) return a.blk < b.blk; // Rotate sends whole blocks: A then B per block (DSP.cpp:312-313)
		return a.row < b.row;
	});
	for (const FrameRec &r : h->h_ring) {
		h->counters[0]++;
		aisgpu_msg m;
		memset(&m, 0, sizeof(m));
		m.stream = r.row >> 1;
		m.channel = (r.row & 1) ? h->cfg.channel_b : h->cfg.channel_a;
		m.nbits = (r.nbits >= 0 && r.nbits <= 1064) ? r.nbits : 0; // Message::setLength (Message.h:288-292)
		m.start_idx = r.start_idx;
		m.end_idx = r.end_idx;
		m.ppm = r.ppm;
		m.chunk = r.chunk;
		float lvl = r.level;
		if ((h->cfg.tag_mode & 1) && lvl != 0.0) lvl = (float)(10.0f * log10((double)lvl)); // AIS.cpp:74-75: the reference resolves to the double log10
		m.level = lvl;
		memcpy(m.data, r.data, 140);
		if (!msg_validate(m.data, m.nbits)) continue; // AIS.cpp:87-93: dropped, siblings were still reset
		build_nmea(m, h->cfg.own_mmsi, &h->seq[m.stream]);
		h->counters[1]++;
		h->counters[(r.row & 1) ? 6 : 5]++;
		h->out_queue.push_back(m);
	}
	return 0;
}

} // namespace

extern "C" {

int aisgpu_abi_version(void) { return AISGPU_ABI_VERSION; }

void aisgpu_default_config(aisgpu_config *cfg) {
	memset(cfg, 0, sizeof(*cfg));
	cfg->struct_size = sizeof(*cfg);
	cfg->model = AISGPU_MODEL_DEFAULT;
	cfg->sample_rate = 1536000;
	cfg->format = AISGPU_FMT_CF32;
	cfg->n_streams = 1;
	cfg->max_chunk_samples = 131072;
	cfg->ps_ema = 1;
	cfg->afc_wide = 1;
	cfg->droop = 1;
	cfg->channel_a = 'A';
	cfg->channel_b = 'B';
	cfg->station = 0;
	cfg->own_mmsi = -1;
	cfg->tag_mode = 3;
	cfg->device = 0;
	cfg->enable_taps = 0;
	cfg->max_frames = 0;
	cfg->host_staging = 1;
	cfg->dsk = 0;
	cfg->fp_ds = 0;
	cfg->dd_train = 0.75f;
	cfg->dd_weight = 0.86f;
}

const char *aisgpu_last_error(aisgpu_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// hooks for the host-only units of the library (host_internal.h)
const aisgpu_config *aisgpu_internal_config(aisgpu_handle *h) { return &h->cfg; }
void aisgpu_internal_set_error(aisgpu_handle *h, const char *msg) { h->err = msg; }

static int create_impl(aisgpu_handle *h) {
	const aisgpu_config &c = h->cfg;
	if (c.model != AISGPU_MODEL_DEFAULT && c.model != AISGPU_MODEL_STANDARD && c.model != AISGPU_MODEL_BASE && c.model != AISGPU_MODEL_V2 &&
		c.model != AISGPU_MODEL_CHALLENGER) {
		h->err = "unknown model kind";
		return AISGPU_EINVAL;
	}
	if (c.format < 0 || c.format > 3 || c.n_streams < 1 || c.max_chunk_samples < 1) {
		h->err = "bad format / n_streams / max_chunk_samples";
		return AISGPU_EINVAL;
	}
	if (int rc = plan_frontend(h)) return rc;
	if (const char *e = getenv("AISGPU_FE_WARPS")) h->fe_warps = atoi(e);
#ifdef AISGPU_FE_ALL_WARP_COUNTS
	if (h->fe_warps != 2 && h->fe_warps != 8) h->fe_warps = 4;
#else
	h->fe_warps = 4;
#endif
	// rows per warp of the decoder kernel, measured per chain: the coherent chain 1; the FM chain 3 with the round-2b front-end shape
	// (bench.py A/B, three pairs: 0.266 / 0.276 / 0.263 ms per step against 0.273 / 0.278 / 0.278 with 6 rows per warp)
	h->dec_rpw = c.model == AISGPU_MODEL_DEFAULT ? 1 : (c.model == AISGPU_MODEL_STANDARD ? 3 : 6);
	if (const char *e = getenv("AISGPU_DEC_RPW")) {
		h->dec_rpw = atoi(e);
		if (h->dec_rpw != 1 && h->dec_rpw != 3) h->dec_rpw = 6;
	}
	if (const char *e = getenv("AISGPU_DECODER")) {
		h->decoder = atoi(e);
		if (h->decoder != 1 && h->decoder != 2) h->decoder = 3;
	}
	if (const char *e = getenv("AISGPU_CF_ROWS")) h->cf_rows = atoi(e) == 8 ? 8 : 4;
	if (c.model == AISGPU_MODEL_CHALLENGER) { // ModelChallenger always demodulates with PhaseSearchEMA (Model.cpp:646-652) and needs the fused kernel's derotated output
		h->cfg.ps_ema = 1;
	}
	if (const char *e = getenv("AISGPU_FE_TILE")) h->fe_tile = atoi(e);
	if (const char *e = getenv("AISGPU_FE_ST")) h->fe_st = atoi(e) ? 1 : 0;
	if (const char *e = getenv("AISGPU_ST_L")) h->st_L = atoi(e);
	if (const char *e = getenv("AISGPU_ST_NB")) h->st_ring = atoi(e) == 3 ? 3 : 5;
	if (const char *e = getenv("AISGPU_ST_KMAX")) h->st_kmax = atoi(e);
	if (const char *e = getenv("AISGPU_FE_CTAS")) h->fe_ctas = std::max(1, atoi(e));
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
		h->err = "no CUDA device (the B200 path has no CPU fallback)";
		return AISGPU_ENODEV;
	}
	if (c.device < 0 || c.device >= ndev) {
		h->err = "CUDA device ordinal out of range";
		return AISGPU_ENODEV;
	}
	CU(cudaSetDevice(c.device));
	// Back-end kernels are small and latency bound; give them priority so that their CTAs are dispatched as soon as the
	// (large-grid) front end of the next submit frees a slot, instead of queueing behind all of its CTAs.
	int prio_lo = 0, prio_hi = 0;
	CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	int prio_mode = 1;
	if (const char *e = getenv("AISGPU_PRIO")) prio_mode = atoi(e); // 0: one priority, 1: back end above front end, 2: the reverse (experiments)
	if (prio_mode == 0) prio_hi = prio_lo;
	else if (prio_mode == 2) std::swap(prio_lo, prio_hi);
	CU(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));
	h->be_streams[0] = h->be_streams[1] = h->stream;
	{
		const char *e = getenv("AISGPU_BE_PIPE");
		// Overlapping the stages of consecutive submits over two back-end streams pays +6 % for the FM chain.  For the coherent chain
		// it is what keeps the step time stable: with one back-end stream the run can lock into a serial pattern (front end c+1 starved
		// while back end c runs, back end c+1 then waiting for it) -- 0.80 ms per step (or 0.90 ms with one stream priority) instead of
		// 0.55 ms at 1024 x 131072 @1536K, both patterns self-sustaining from the first submits on (profiles/r2_sweeps.jsonl, probe12/13).
		const bool pipe = (e ? atoi(e) != 0 : (c.model == AISGPU_MODEL_STANDARD || c.model == AISGPU_MODEL_DEFAULT || c.model == AISGPU_MODEL_CHALLENGER)) && !c.enable_taps &&
						  c.model != AISGPU_MODEL_BASE && c.model != AISGPU_MODEL_V2;
		if (pipe) CU(cudaStreamCreateWithPriority(&h->be_streams[1], cudaStreamNonBlocking, prio_hi));
	}
	h->bs = h->stream;
	for (int st = 0; st < aisgpu_handle::NSTAGE; st++)
		for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&h->ev_stage[st][i], cudaEventDisableTiming));
	for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&h->ev_ec_read[i], cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
	CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithPriority(&h->fe_stream, cudaStreamNonBlocking, prio_lo));
	for (int i = 0; i < aisgpu_handle::NC; i++) {
		CU(cudaEventCreateWithFlags(&h->ev_fe_done[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&h->ev_be_done[i], cudaEventDisableTiming));
	}
	for (int i = 0; i < 3; i++) {
		CU(cudaEventCreateWithFlags(&h->ev_rot[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&h->ev_k1[i], cudaEventDisableTiming));
	}
	for (int i = 0; i < aisgpu_handle::NEV; i++) {
		CU(cudaEventCreate(&h->ev_fe0s[i]));
		CU(cudaEventCreate(&h->ev_fe1s[i]));
	}
	for (int i = 0; i < 2; i++) {
		CU(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
	}
	const int B = c.n_streams, k = h->k;
	const int q = outer_granule(h);
	const int maxN = (c.max_chunk_samples + q - 1) / q * q;
	h->cfg.max_chunk_samples = maxN;
	h->rows = 2 * B;
	h->obps = bytes_per_sample(c.format);
	h->bps = bytes_per_sample(h->in_fmt);
	h->inner_max = h->pre == 1 ? (maxN >> h->kA) : (h->pre >= 2 ? h->blk : maxN);
	h->max_n48 = h->inner_max >> (k + 1);
	h->seq.assign(B, 0);
	if (h->pre) { // resampler pre-stage: raw-format history, decimated stream, schedule tables, ring of reference blocks
		const int tl = h->pre == 2 ? 32 : h->PA;
		for (int i = 0; i < 2; i++) {
			if (int rc = dalloc(h, &h->d_ptail[i], (size_t)B * tl * h->obps)) return rc;
			if (c.format == AISGPU_FMT_CU8) CU(cudaMemsetAsync(h->d_ptail[i], 0x80, (size_t)B * tl * h->obps, h->stream));
		}
		if (h->pre == 1 || h->pre == 3 || h->pre == 4) {
			const int Lmax = maxN >> h->kA;
			h->d0_stride = (Lmax + 4 + 1) & ~1LL;
			if (int rc = dalloc(h, &h->d_D0, (size_t)B * h->d0_stride)) return rc;
			memset(&h->fe_pre, 0, sizeof(h->fe_pre));
			if (h->pre != 4) {
				if (int rc = dalloc(h, &h->d_us_src, (size_t)2 * Lmax + 8)) return rc;
				if (int rc = dalloc(h, &h->d_us_alpha, (size_t)2 * Lmax + 8)) return rc;
				h->s_stride = 4LL * Lmax;
			}
		}
		if (h->pre >= 2) {
			const int cap96 = ((2 * maxN / 3 + 1 + h->blk + h->blk - 1) / h->blk + 1) * h->blk; // behind Upsample up to 2x the samples
			CU(set_taps_bh28_3(H_TAPS_BH28_3));
			if (h->pre == 2 || h->pre == 4) { // DownsampleKFilter writes the 96 kS/s ring directly
				h->s_cap = cap96;
				h->s_stride = cap96;
			}
			else {
				h->s2_cap = cap96;
				h->s2_stride = cap96;
				if (int rc = dalloc(h, &h->d_S2, (size_t)B * h->s2_stride)) return rc;
			}
			if (h->pre != 2)
				for (int i = 0; i < 2; i++)
					if (int rc = dalloc(h, &h->d_ptail2[i], (size_t)B * 32)) return rc;
		}
		if (int rc = dalloc(h, &h->d_S, (size_t)B * h->s_stride)) return rc;
	}
	for (int i = 0; i < 2; i++) {
		if (int rc = dalloc(h, &h->d_tail[i], (size_t)B * h->P * h->bps)) return rc;
		if (h->in_fmt == AISGPU_FMT_CU8 && !h->fp_ds) // the reference's zero initial filter state is byte value 128 in CU8 (0 for the unbiased integer pipeline)
			CU(cudaMemsetAsync(h->d_tail[i], 0x80, (size_t)B * h->P * h->bps, h->stream));
		if (int rc = dalloc(h, &h->d_fir_hist[i], (size_t)h->rows * 16)) return rc;
	}
	for (int i = 0; i < 3; i++)
		if (int rc = dalloc(h, &h->d_rot[i], (size_t)h->P96 + (h->inner_max >> k) + 8)) return rc;
	if (int rc = dalloc(h, &h->d_rot_state, 4)) return rc;
	{
		float2 one = make_float2(1.0f, 0.0f);
		CU(cudaMemcpyAsync(h->d_rot_state, &one, sizeof(one), cudaMemcpyHostToDevice, h->stream)); // after the memset on the same stream
		h->mult = polar1((float)(PI_F * 25000.0 / 48000.0)); // Model.cpp:31
	}
	h->c_stride = (HC + h->max_n48 + 8 + 1) & ~1LL;
	for (int i = 0; i < 2; i++)
		if (int rc = dalloc(h, &h->d_C2[i], (size_t)h->rows * h->c_stride)) return rc;
	if (int rc = dalloc(h, &h->d_C2[2], (size_t)h->rows * h->c_stride)) return rc;
	const int nEmax = HC + h->max_n48;
	h->e_stride = (HE + nEmax + 8 + 1) & ~1LL;
	h->r_stride = nEmax;
	const bool coherent = c.model == AISGPU_MODEL_DEFAULT || c.model == AISGPU_MODEL_CHALLENGER; // the CGF / FIR17 / PhaseSearch chain
	const int ndec = c.model == AISGPU_MODEL_V2 ? 6 : (c.model == AISGPU_MODEL_CHALLENGER ? 10 : 5); // decoders per row
	if (int rc = dalloc(h, &h->d_dec, (size_t)h->rows * ndec)) return rc;
	if (int rc = dalloc(h, &h->d_dec_data, (size_t)h->rows * ndec * DEC_WORDS)) return rc;
	if (c.model == AISGPU_MODEL_V2) {
		h->c_hist = V2_BLK; // Engine::raw starts as a block of zeros that is decoded when the first real block has arrived (V2Engine.cpp:274-277, 379-395)
		if (int rc = dalloc(h, &h->d_v2, (size_t)h->rows)) return rc;
		if (int rc = dalloc(h, &h->d_omega, CGF_N)) return rc;
		{
			std::vector<V2State> init(h->rows);
			memset(init.data(), 0, init.size() * sizeof(V2State));
			for (auto &v : init) {
				v.fo_rot = make_float2(1.0f, 0.0f);
				v.fm_prev = make_float2(1.0f, 0.0f);
			}
			std::vector<float2> om(CGF_N);
			for (int s = 0; s < CGF_N; s++) om[s] = polar1((float)(-2.0 * PI_F) * (float)s / (float)CGF_N);
			CU(cudaMemcpyAsync(h->d_v2, init.data(), init.size() * sizeof(V2State), cudaMemcpyHostToDevice, h->stream));
			CU(cudaMemcpyAsync(h->d_omega, om.data(), om.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
			CU(cudaStreamSynchronize(h->stream));
			CU(v2_init(H_TAPS_COHERENT, H_TAPS_RECEIVER, om.data()));
		}
		if (c.enable_taps) {
			if (int rc = dalloc(h, &h->d_tap_cgf, (size_t)h->rows * h->r_stride)) return rc;
			if (int rc = dalloc(h, &h->d_tap_coh, (size_t)h->rows * h->r_stride)) return rc;
			if (int rc = dalloc(h, &h->d_tap_fm, (size_t)h->rows * h->r_stride)) return rc;
		}
	}
	else if (coherent) {
		h->c_hist = 0;
		for (int i = 0; i < 2; i++) {
			if (int rc = dalloc(h, &h->d_stepidx2[i], (size_t)h->rows * (nEmax / CGF_N + 1))) return rc;
		}
		if (int rc = dalloc(h, &h->d_cgf_rot, (size_t)h->rows)) return rc;
		for (int i = 0; i < 2; i++)
			if (int rc = dalloc(h, &h->d_Ec2[i], (size_t)h->rows * h->e_stride)) return rc;
		if (int rc = dalloc(h, &h->d_ps, (size_t)h->rows * 5)) return rc;
		h->dwords = (nEmax / 5 + 1 + K3_TS - 1) / K3_TS + 1;
		for (int i = 0; i < 2; i++) {
			if (int rc = dalloc(h, &h->d_dbits2[i], (size_t)h->rows * 5 * h->dwords)) return rc;
			if (int rc = dalloc(h, &h->d_lvl2[i], (size_t)h->rows * h->dwords * K3_TS)) return rc;
		}
		if (!c.ps_ema)
			if (int rc = dalloc(h, &h->d_ps_mem, (size_t)h->rows * 5 * 16 * 12)) return rc;
		if (c.model == AISGPU_MODEL_CHALLENGER) { // FM branch on the derotated samples (Model.cpp:637-639)
			h->ed_stride = (HD + nEmax + 8 + 1) & ~1LL;
			if (int rc = dalloc(h, &h->d_Ed, (size_t)h->rows * h->ed_stride)) return rc;
			if (int rc = dalloc(h, &h->d_Ef2[0], (size_t)h->rows * h->e_stride)) return rc;
			for (int i = 0; i < 2; i++)
				if (int rc = dalloc(h, &h->d_dbitsF[i], (size_t)h->rows * 5 * h->dwords)) return rc;
			if (int rc = dalloc(h, &h->d_lvl_prev, (size_t)2 * h->rows)) return rc;
		}
		if (int rc = dalloc(h, &h->d_steptab, CGF_NIDX)) return rc;
		if (int rc = dalloc(h, &h->d_ppmtab, CGF_NIDX)) return rc;
		if (int rc = dalloc(h, &h->d_omega, CGF_N)) return rc;
		std::vector<float2> one(h->rows, make_float2(1.0f, 0.0f));
		CU(cudaMemcpyAsync(h->d_cgf_rot, one.data(), one.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
		std::vector<float2> om(CGF_N), st(CGF_NIDX);
		std::vector<float> pp(CGF_NIDX);
		for (int s = 0; s < CGF_N; s++) om[s] = polar1((float)(-2.0 * PI_F) * (float)s / (float)CGF_N); // FFT.h:81-83
		for (int idx = 0; idx < CGF_NIDX; idx++) { // DSP.cpp:453,457-458,466
			float fz = -1;
			if (idx != CGF_IDX_NONE) {
				int i = idx - CGF_IDX_OFFSET;
				fz = (CGF_N / 2 - (i + 102 / 2.0f));
			}
			float f = fz / 2.0f / CGF_N;
			st[idx] = polar1((float)(f * 2 * PI_F));
			pp[idx] = f * 48000.0f / 162.0f;
		}
		CU(cudaMemcpyAsync(h->d_omega, om.data(), om.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(h->d_steptab, st.data(), st.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
		CU(cudaMemcpyAsync(h->d_ppmtab, pp.data(), pp.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
		CU(cudaStreamSynchronize(h->stream)); // host vectors go out of scope
		if (c.enable_taps)
			if (int rc = dalloc(h, &h->d_tap_cgf, (size_t)h->rows * h->r_stride)) return rc;
	}
	else {
		h->c_hist = FIRF_T;
		if (int rc = dalloc(h, &h->d_Ef2[0], (size_t)h->rows * h->e_stride)) return rc;
		h->dwords = (nEmax / 5 + 2 + K3_TS - 1) / K3_TS + 1;
		for (int i = 0; i < 2; i++)
			if (int rc = dalloc(h, &h->d_dbits2[i], (size_t)h->rows * 5 * h->dwords)) return rc;
		if (c.model == AISGPU_MODEL_BASE) {
			if (int rc = dalloc(h, &h->d_pll, (size_t)h->rows)) return rc;
			std::vector<PllState> pl(h->rows);
			for (auto &x : pl) { x.prev = 0; x.pll = 0.0f; x.fast = 1; }
			CU(cudaMemcpyAsync(h->d_pll, pl.data(), pl.size() * sizeof(PllState), cudaMemcpyHostToDevice, h->stream));
			CU(cudaStreamSynchronize(h->stream));
		}
		if (c.enable_taps) {
			if (int rc = dalloc(h, &h->d_tap_fm, (size_t)h->rows * h->r_stride)) return rc;
			if (int rc = dalloc(h, &h->d_tap_cnt, (size_t)h->rows)) return rc;
		}
	}
	if (c.enable_taps)
		if (int rc = dalloc(h, &h->d_tap_dec, (size_t)h->rows * 5 * (nEmax / 5 + 2))) return rc;
	if (getenv("AISGPU_DEBUG"))
		if (int rc = dalloc(h, &h->d_dbg, (size_t)h->rows * 4)) return rc;
	{
		std::vector<float2> om(CGF_N / 2);
		for (int s = 0; s < CGF_N / 2; s++) om[s] = polar1((float)(-2.0 * PI_F) * (float)s / (float)CGF_N); // FFT.h:81-83
		CU(cgf_init(H_TAPS_COHERENT, om.data()));
	}
	CU(fm_init(H_TAPS_RECEIVER));
	{
		uint32_t ab[35] = { 0 };
		const int pos[] = { 30, 62, 96, 168, 184, 192, 336, 385, 448, MAX_FRAME_BITS }; // AIS.cpp:111-142, AIS.h:172
		for (int q : pos) ab[q >> 5] |= 1u << (q & 31);
		CU(sym_init(H_PS_COS, H_PS_SIN, ab));
	}
	h->ring_cap = c.max_frames > 0 ? c.max_frames : std::max(4096, B * 64);
	if (int rc = dalloc(h, &h->d_ring, (size_t)h->ring_cap)) return rc;
	if (int rc = dalloc(h, &h->d_ring_head, 1)) return rc;
	if (int rc = dalloc(h, &h->d_counts, 16)) return rc;
	if (cudaMallocHost((void **)&h->pin_head, aisgpu_handle::NT * sizeof(unsigned long long)) != cudaSuccess) {
		h->err = "out of pinned host memory";
		(void)cudaGetLastError();
		return AISGPU_ENOMEM;
	}
	memset(h->pin_head, 0, aisgpu_handle::NT * sizeof(unsigned long long));
	for (int i = 0; i < aisgpu_handle::NT; i++) CU(cudaEventCreateWithFlags(&h->ev_ticket[i], cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&h->ev_mark, cudaEventDisableTiming));
	if (h->pre == 1 || h->pre == 3) {
		h->us_cap = 2 * (maxN >> h->kA) + 8; // Upsample never more than doubles the rate (bucket / 2 < rate)
		for (int i = 0; i < 2; i++) {
			if (cudaMallocHost((void **)&h->pin_us_src[i], (size_t)h->us_cap * sizeof(int)) != cudaSuccess ||
				cudaMallocHost((void **)&h->pin_us_alpha[i], (size_t)h->us_cap * sizeof(float)) != cudaSuccess) {
				h->err = "out of pinned host memory";
				(void)cudaGetLastError();
				return AISGPU_ENOMEM;
			}
			CU(cudaEventCreateWithFlags(&h->ev_us[i], cudaEventDisableTiming));
		}
	}
	if (c.host_staging) // the H2D staging buffers of aisgpu_submit / _v / _async (otherwise allocated by the first host submit)
		for (int i = 0; i < 2; i++)
			if (int rc = dalloc(h, &h->d_in[i], (size_t)B * maxN * h->obps)) return rc;
	memset(&h->fe, 0, sizeof(h->fe));
	CU(cudaStreamSynchronize(h->stream));
	return 0;
}

int aisgpu_create(const aisgpu_config *cfg, aisgpu_handle **out) {
	if (!cfg || !out || cfg->struct_size != sizeof(aisgpu_config)) {
		g_create_error = "aisgpu_create: null argument or struct_size mismatch";
		return AISGPU_EINVAL;
	}
	aisgpu_handle *h = new aisgpu_handle();
	h->cfg = *cfg;
	int rc = create_impl(h);
	if (rc) {
		g_create_error = h->err;
		aisgpu_destroy(h);
		*out = nullptr;
		return rc;
	}
	*out = h;
	return 0;
}

// a CUDA failure (or an internal inconsistency) inside a submit leaves carried state half-advanced: poison the handle
static int poison(aisgpu_handle *h, int rc) {
	if (rc != 0 && rc != AISGPU_EINVAL && !h->poisoned) h->poisoned = rc;
	return rc;
}
#define ENTER(h)                                  \
	do {                                          \
		if (!(h)) return AISGPU_EINVAL;           \
		if ((h)->poisoned) return (h)->poisoned;  \
		CU(cudaSetDevice((h)->cfg.device));       \
	} while (0)

int aisgpu_submit_device(aisgpu_handle *h, const void *dev_samples, int64_t stride_samples, int n_samples) {
	ENTER(h);
	if (!dev_samples) return AISGPU_EINVAL;
	if (stride_samples < n_samples || (stride_samples & 1)) {
		h->err = "stride_samples must be even and >= n_samples";
		return AISGPU_EINVAL;
	}
	if (int rc = check_outer(h, n_samples)) return rc; // all argument checks come before any state is touched
	return poison(h, submit_outer(h, dev_samples, stride_samples, n_samples));
}

// Host submits: the batch goes through one of two device staging buffers (copy stream), the front end of the submit
// before last being the previous reader of that buffer.  `ptrs` != nullptr: one host pointer per stream.
static int submit_host(aisgpu_handle *h, const void *host_samples, const void *const *ptrs, int n_samples, bool wait_copy, int64_t *ticket) {
	if (int rc = check_outer(h, n_samples)) return rc;
	const size_t row_bytes = (size_t)n_samples * h->obps;
	const int cur = h->in_cur;
	if (!h->d_in[cur]) {
		if (int rc = dalloc(h, &h->d_in[cur], (size_t)h->cfg.n_streams * h->cfg.max_chunk_samples * h->obps)) return rc;
		CU(cudaStreamSynchronize(h->stream)); // dalloc clears on h->stream
	}
	if (h->in_used[cur]) CU(cudaStreamWaitEvent(h->copy_stream, h->ev_done[cur], 0));
	if (ptrs) {
		for (int s = 0; s < h->cfg.n_streams; s++) {
			if (!ptrs[s]) {
				h->err = "null stream pointer";
				return AISGPU_EINVAL;
			}
		}
		for (int s = 0; s < h->cfg.n_streams; s++)
			CU(cudaMemcpyAsync(h->d_in[cur] + (size_t)s * row_bytes, ptrs[s], row_bytes, cudaMemcpyHostToDevice, h->copy_stream));
	}
	else CU(cudaMemcpyAsync(h->d_in[cur], host_samples, row_bytes * h->cfg.n_streams, cudaMemcpyHostToDevice, h->copy_stream));
	CU(cudaEventRecord(h->ev_copy[cur], h->copy_stream));
	CU(cudaStreamWaitEvent(h->fe_stream, h->ev_copy[cur], 0));
	if (ticket) *ticket = (int64_t)h->counters[3];
	if (int rc = submit_outer(h, h->d_in[cur], n_samples, n_samples)) return rc;
	CU(cudaEventRecord(h->ev_done[cur], h->fe_stream)); // the front end is the only reader of the staging buffer
	h->in_used[cur] = true;
	h->in_cur ^= 1;
	// aisgpu_submit / _v: the caller's buffer is only borrowed for the call (Stream.h:41 semantics): wait for the copy, not
	// for the kernels
	if (wait_copy) CU(cudaEventSynchronize(h->ev_copy[cur]));
	return 0;
}

int aisgpu_submit(aisgpu_handle *h, const void *host_samples, int n_samples) {
	ENTER(h);
	if (!host_samples) return AISGPU_EINVAL;
	return poison(h, submit_host(h, host_samples, nullptr, n_samples, true, nullptr));
}

int aisgpu_submit_v(aisgpu_handle *h, const void *const *stream_ptrs, int n_samples) {
	ENTER(h);
	if (!stream_ptrs) return AISGPU_EINVAL;
	return poison(h, submit_host(h, nullptr, stream_ptrs, n_samples, true, nullptr));
}

int aisgpu_submit_async(aisgpu_handle *h, const void *host_samples, int n_samples, int64_t *ticket) {
	ENTER(h);
	if (!host_samples) return AISGPU_EINVAL;
	return poison(h, submit_host(h, host_samples, nullptr, n_samples, false, ticket));
}

int aisgpu_sync(aisgpu_handle *h) {
	ENTER(h);
	CU(cudaStreamSynchronize(h->copy_stream));
	CU(cudaStreamSynchronize(h->fe_stream));
	if (int rc = sync_backend(h)) return rc;
	return 0;
}

int aisgpu_poll_upto(aisgpu_handle *h, int64_t ticket, aisgpu_msg *out, int max, int *n) {
	if (!n || (max > 0 && !out)) return AISGPU_EINVAL;
	ENTER(h);
	if (h->out_pos >= h->out_queue.size()) {
		h->out_queue.clear();
		h->out_pos = 0;
		if (int rc = drain_ring(h, (long long)ticket)) return rc;
	}
	int k = 0;
	while (k < max && h->out_pos < h->out_queue.size()) out[k++] = h->out_queue[h->out_pos++];
	*n = k;
	if (h->overflow_pending) { // reported once per loss, together with the frames that survived
		h->overflow_pending = false;
		h->err = "frame ring overflow: frames were dropped (see counters[4]); raise aisgpu_config.max_frames or poll more often";
		return AISGPU_EOVERFLOW;
	}
	return 0;
}

int aisgpu_poll(aisgpu_handle *h, aisgpu_msg *out, int max, int *n) { return aisgpu_poll_upto(h, -1, out, max, n); }

int aisgpu_tap(aisgpu_handle *h, int tap, int stream, int channel, void *dst, size_t dst_bytes, size_t *n_out) {
	if (!h || !n_out || stream < 0 || stream >= h->cfg.n_streams || channel < 0 || channel > 9) return AISGPU_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	CU(cudaStreamSynchronize(h->fe_stream));
	if (int rc = sync_backend(h)) return rc;
	const int row = stream * 2 + (channel & 1);
	const void *src = nullptr;
	size_t n = 0, esz = 8;
	switch (tap) {
	case AISGPU_TAP_C:
		src = h->d_C2[h->c_last] + (long long)row * h->c_stride + HC; // note: valid until the next submit only for [0, n48)
		n = h->last_n48;
		break;
	case AISGPU_TAP_CGF:
		if (!h->d_tap_cgf) { h->err = "taps not enabled or not a ModelDefault engine"; return AISGPU_EINVAL; }
		src = h->d_tap_cgf + (long long)row * h->r_stride;
		n = h->last_nE;
		break;
	case AISGPU_TAP_FIR:
		if (h->cfg.model == AISGPU_MODEL_V2) {
			if (!h->d_tap_coh) { h->err = "taps not enabled"; return AISGPU_EINVAL; }
			src = h->d_tap_coh + (long long)row * h->r_stride;
		}
		else if (h->cfg.model == AISGPU_MODEL_DEFAULT) src = h->d_Ec2[h->ec_last] + (long long)row * h->e_stride + HE;
		else {
			if (!h->cfg.enable_taps && h->cfg.model != AISGPU_MODEL_BASE) { h->err = "taps not enabled"; return AISGPU_EINVAL; } // the FIR37 output is not stored then
			src = h->d_Ef2[0] + (long long)row * h->e_stride + HE;
			esz = 4;
		}
		n = h->last_nE;
		break;
	case AISGPU_TAP_ROT:
		CU(cudaStreamSynchronize(h->side_stream));
		src = h->d_rot[h->rot_cur] + h->P96;
		n = h->rot_n96[h->rot_cur];
		break;
	case 4: { // decoder input of sampling phase (channel / 2): channel = ch + 2 * phase
		if (!h->d_tap_dec) { h->err = "taps not enabled"; return AISGPU_EINVAL; }
		const int phase = channel >> 1;
		esz = 4;
		if (h->cfg.model == AISGPU_MODEL_BASE) {
			int cnt = 0;
			CU(cudaMemcpy(&cnt, h->d_tap_cnt + row, sizeof(int), cudaMemcpyDeviceToHost));
			src = h->d_tap_dec + (long long)row * h->last_nE;
			n = cnt;
		}
		else {
			src = h->d_tap_dec + (long long)(row * 5 + phase) * h->last_nsym;
			n = h->last_nsym;
			if (h->cfg.model == AISGPU_MODEL_STANDARD) { // samples of this phase among the last submit's [a0, a1)
				const long long a1 = h->e_abs, a0 = a1 - h->last_nE;
				n = 0;
				for (long long a = a0; a < a1; a++) n += (a % 5) == phase;
			}
		}
		break;
	}
	case AISGPU_TAP_PRE:
	case AISGPU_TAP_PRE2: { // what the resampler in front of the decimation chain produced in the last submit (stream, channel ignored)
		const bool second = tap == AISGPU_TAP_PRE2;
		const float2 *ring = second ? h->d_S2 : h->d_S;
		const long long a = second ? h->pre2_tap0 : h->pre_tap0, b = second ? h->pre2_tap1 : h->pre_tap1;
		const int cap = second ? h->s2_cap : h->s_cap;
		const long long st = second ? h->s2_stride : h->s_stride;
		if (!ring || cap <= 0) { h->err = "this rate has no resampler stage"; return AISGPU_EINVAL; }
		size_t cnt = (size_t)std::min<long long>(b - a, cap);
		if (dst_bytes < cnt * 8) cnt = dst_bytes / 8;
		for (size_t i = 0; i < cnt;) { // the ring may wrap
			const size_t off = (size_t)((a + (long long)i) % cap), run = std::min(cnt - i, (size_t)cap - off);
			if (dst) CU(cudaMemcpy((float2 *)dst + i, ring + (long long)stream * st + off, run * 8, cudaMemcpyDeviceToHost));
			i += run;
		}
		*n_out = cnt;
		return 0;
	}
	case 6: // debug counters of the decoder kernel, 4 x int64 per row (stream/channel ignored, all rows)
		if (!h->d_dbg) { h->err = "AISGPU_DEBUG not set"; return AISGPU_EINVAL; }
		src = h->d_dbg;
		n = (size_t)h->rows * 4;
		break;
	case 5:
		if (!h->d_tap_fm) { h->err = "taps not enabled or not an FM engine"; return AISGPU_EINVAL; }
		src = h->d_tap_fm + (long long)row * h->r_stride;
		n = h->last_nE;
		esz = 4;
		break;
	default:
		h->err = "unknown tap";
		return AISGPU_EINVAL;
	}
	if (dst_bytes < n * esz) n = dst_bytes / esz;
	if (n && dst) CU(cudaMemcpy(dst, src, n * esz, cudaMemcpyDeviceToHost));
	*n_out = n;
	return 0;
}

int aisgpu_counters(aisgpu_handle *h, uint64_t counters[8]) {
	if (!h || !counters) return AISGPU_EINVAL;
	memcpy(counters, h->counters, sizeof(h->counters));
	return 0;
}

// ---- NCCL, resolved at run time: the counters are the only thing that ever crosses NVLink ----
namespace {
struct nccl_uid { char internal[128]; };
struct NcclApi {
	void *lib = nullptr;
	int (*GetUniqueId)(nccl_uid *) = nullptr;
	int (*CommInitRank)(void **, int, nccl_uid, int) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
	int (*CommDestroy)(void *) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
};
NcclApi *nccl_api(std::string &err) {
	static NcclApi api;
	static bool tried = false;
	if (!tried) {
		tried = true;
		const char *names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char *nm : names)
			if ((api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
		if (api.lib) {
			api.GetUniqueId = (int (*)(nccl_uid *))dlsym(api.lib, "ncclGetUniqueId");
			api.CommInitRank = (int (*)(void **, int, nccl_uid, int))dlsym(api.lib, "ncclCommInitRank");
			api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
			api.CommDestroy = (int (*)(void *))dlsym(api.lib, "ncclCommDestroy");
			api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
			if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
				dlclose(api.lib);
				api.lib = nullptr;
			}
		}
	}
	if (!api.lib) {
		err = "NCCL not found (dlopen libnccl.so.2)";
		return nullptr;
	}
	return &api;
}
} // namespace

int aisgpu_nccl_unique_id(void *id128) {
	if (!id128) return AISGPU_EINVAL;
	NcclApi *api = nccl_api(g_create_error);
	if (!api) return AISGPU_ENODEV;
	nccl_uid id;
	if (int rc = api->GetUniqueId(&id)) {
		g_create_error = std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error");
		return AISGPU_ECUDA;
	}
	memcpy(id128, &id, sizeof(id));
	return 0;
}

int aisgpu_comm_init(aisgpu_handle *h, const void *id128, int n_ranks, int rank) {
	if (!h || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return AISGPU_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	NcclApi *api = nccl_api(h->err);
	if (!api) return AISGPU_ENODEV;
	if (h->nccl_comm) {
		api->CommDestroy(h->nccl_comm);
		h->nccl_comm = nullptr;
	}
	nccl_uid id;
	memcpy(&id, id128, sizeof(id));
	if (int rc = api->CommInitRank(&h->nccl_comm, n_ranks, id, rank)) {
		h->err = std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error");
		h->nccl_comm = nullptr;
		return AISGPU_ECUDA;
	}
	return 0;
}

int aisgpu_allreduce_counts(aisgpu_handle *h, uint64_t totals[8]) {
	if (!h || !totals) return AISGPU_EINVAL;
	CU(cudaSetDevice(h->cfg.device));
	if (!h->nccl_comm) { // a single engine is its own job
		memcpy(totals, h->counters, sizeof(h->counters));
		return 0;
	}
	NcclApi *api = nccl_api(h->err);
	if (!api) return AISGPU_ENODEV;
	CU(cudaMemcpyAsync(h->d_counts, h->counters, 8 * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
	if (int rc = api->AllReduce(h->d_counts, h->d_counts + 8, 8, /*ncclUint64*/ 5, /*ncclSum*/ 0, h->nccl_comm, h->stream)) {
		h->err = std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error");
		return AISGPU_ECUDA;
	}
	CU(cudaMemcpyAsync(totals, h->d_counts + 8, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return 0;
}

void *aisgpu_cuda_stream(aisgpu_handle *h) { return h ? (void *)h->stream : nullptr; }

int aisgpu_join(aisgpu_handle *h) {
	if (!h) return AISGPU_EINVAL;
	if (h->be_streams[1] != h->be_streams[0]) {
		CU(cudaEventRecord(h->ev_join, h->be_streams[1]));
		CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
	}
	return 0;
}

float aisgpu_last_frontend_ms(aisgpu_handle *h) {
	float ms = -1.0f;
	int n = 0;
	if (aisgpu_frontend_times(h, &ms, 1, &n) || n != 1) return -1.0f;
	return ms;
}

int aisgpu_frontend_times(aisgpu_handle *h, float *ms_out, int max, int *n) {
	if (!h || !ms_out || !n) return AISGPU_EINVAL;
	*n = 0;
	if (!h->fe_timed) return 0;
	CU(cudaStreamSynchronize(h->fe_stream));
	if (int rc = sync_backend(h)) return rc;
	long long cnt = std::min<long long>(std::min<long long>(h->chunk, aisgpu_handle::NEV), max);
	for (long long i = 0; i < cnt; i++) { // newest first
		const int evi = (int)((h->chunk - 1 - i) % aisgpu_handle::NEV);
		CU(cudaEventElapsedTime(&ms_out[i], h->ev_fe0s[evi], h->ev_fe1s[evi]));
	}
	*n = (int)cnt;
	return 0;
}

int aisgpu_last_launches(aisgpu_handle *h) { return h ? h->last_launches : 0; }

int aisgpu_chunk_granule(const aisgpu_config *cfg) {
	if (!cfg || cfg->struct_size != sizeof(aisgpu_config)) {
		g_create_error = "aisgpu_chunk_granule: null argument or struct_size mismatch";
		return AISGPU_EINVAL;
	}
	aisgpu_handle tmp;
	tmp.cfg = *cfg;
	if (int rc = plan_frontend(&tmp)) {
		g_create_error = tmp.err;
		return rc;
	}
	return outer_granule(&tmp);
}

int aisgpu_validate(const uint8_t *data, int nbits) {
	if (!data || nbits < 0) return 0;
	return msg_validate(data, nbits) ? 1 : 0;
}

int aisgpu_build_nmea(aisgpu_msg *m, int own_mmsi, int *seq) {
	if (!m || !seq || m->nbits < 0 || m->nbits > 1064 || *seq < 0 || *seq > 9) return AISGPU_EINVAL;
	build_nmea(*m, own_mmsi, seq);
	return 0;
}

void aisgpu_destroy(aisgpu_handle *h) {
	if (!h) return;
	if (h->fe_stream) cudaStreamSynchronize(h->fe_stream);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->be_streams[1] && h->be_streams[1] != h->stream) { cudaStreamSynchronize(h->be_streams[1]); cudaStreamDestroy(h->be_streams[1]); }
	for (int st = 0; st < aisgpu_handle::NSTAGE; st++)
		for (int i = 0; i < 2; i++)
			if (h->ev_stage[st][i]) cudaEventDestroy(h->ev_stage[st][i]);
	for (int i = 0; i < 2; i++)
		if (h->ev_ec_read[i]) cudaEventDestroy(h->ev_ec_read[i]);
	if (h->ev_join) cudaEventDestroy(h->ev_join);
	if (h->nccl_comm) {
		std::string e;
		if (NcclApi *api = nccl_api(e)) api->CommDestroy(h->nccl_comm);
	}
	if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
	for (int i = 0; i < aisgpu_handle::NT; i++)
		if (h->ev_ticket[i]) cudaEventDestroy(h->ev_ticket[i]);
	if (h->ev_mark) cudaEventDestroy(h->ev_mark);
	if (h->pin_head) cudaFreeHost(h->pin_head);
	for (int i = 0; i < 2; i++) {
		if (h->pin_us_src[i]) cudaFreeHost(h->pin_us_src[i]);
		if (h->pin_us_alpha[i]) cudaFreeHost(h->pin_us_alpha[i]);
		if (h->ev_us[i]) cudaEventDestroy(h->ev_us[i]);
	}
	void *ptrs[] = { h->d_ptail[0], h->d_ptail[1], h->d_ptail2[0], h->d_ptail2[1], h->d_S2, h->d_D0, h->d_S, h->d_us_src, h->d_us_alpha, h->d_in[0], h->d_in[1], h->d_tail[0], h->d_tail[1], h->d_rot[0], h->d_rot[1], h->d_rot[2], h->d_rot_state, h->d_C2[0], h->d_C2[1], h->d_C2[2],
					 h->d_steptab, h->d_omega, h->d_cgf_rot, h->d_stepidx2[0], h->d_stepidx2[1], h->d_ppmtab, h->d_fir_hist[0], h->d_fir_hist[1], h->d_tap_cgf, h->d_Ec2[0], h->d_Ec2[1],
					 h->d_Ef2[0], h->d_ps, h->d_ps_mem, h->d_dbits2[0], h->d_dbits2[1], h->d_lvl2[0], h->d_lvl2[1], h->d_dbg, h->d_dec, h->d_dec_data, h->d_pll, h->d_tap_dec, h->d_tap_fm, h->d_tap_cnt, h->d_v2, h->d_tap_coh, h->d_Ed, h->d_dbitsF[0], h->d_dbitsF[1], h->d_lvl_prev, h->d_ring,
					 h->d_ring_head, h->d_counts };
	for (void *p : ptrs)
		if (p) cudaFree(p);
	for (int i = 0; i < aisgpu_handle::NEV; i++) {
		if (h->ev_fe0s[i]) cudaEventDestroy(h->ev_fe0s[i]);
		if (h->ev_fe1s[i]) cudaEventDestroy(h->ev_fe1s[i]);
	}
	for (int i = 0; i < 2; i++) {
		if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
		if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
	}
	for (int i = 0; i < 3; i++) {
		if (h->ev_rot[i]) cudaEventDestroy(h->ev_rot[i]);
		if (h->ev_k1[i]) cudaEventDestroy(h->ev_k1[i]);
	}
	if (h->side_stream) { cudaStreamSynchronize(h->side_stream); cudaStreamDestroy(h->side_stream); }
	for (int i = 0; i < aisgpu_handle::NC; i++) {
		if (h->ev_fe_done[i]) cudaEventDestroy(h->ev_fe_done[i]);
		if (h->ev_be_done[i]) cudaEventDestroy(h->ev_be_done[i]);
	}
	if (h->fe_stream) cudaStreamDestroy(h->fe_stream);
	if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
}

} // extern "C"
