// Test program for ais-catcher_b200/host/ModelGPU.h: the adapter inside the reference's own block graph.
//
// Links the UNMODIFIED reference objects (oracle/_ref/strict/*.o, compiled from /root/reference by oracle/Makefile)
// with libaisgpu.so and wires  MemDevice --Connection<RAW>--> AIS::ModelGPU --StreamOut<Message>--> sink , i.e. what
// Receiver::setupModel + Engine::run do for a CPU model (reference Source/Application/Receiver.cpp:199-244).
//
//   adapter_test <file> <format CU8|CF32> <sample_rate> <block_samples> <model 0|1|2> [cpu]
//
// prints one line per message: channel|nbits|start|end|level-bits|ppm-bits|sentence[ sentence...]
// With the trailing "cpu" the same graph is built with the reference's CPU model instead (for an A/B in one binary).
// Exit code 3 = the adapter reported a run-time failure through Error() + StopRequest() (e.g. no GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "Device.h"
#include "Model.h"
#include "ModelGPU.h"

static int g_stop_requests = 0;
void StopRequest() { g_stop_requests++; } // Source/Library/Common.h:72 -- the application normally defines it

namespace {
struct MemDevice : public Device::Device {
	void push(void *p, int bytes, Format f) {
		RAW r{f, p, bytes};
		Send(&r, 1, tag);
	}
};
struct Sink : public StreamIn<AIS::Message> {
	long count = 0;
	void Receive(const AIS::Message *m, int len, TAG &tag) override {
		for (int i = 0; i < len; i++) {
			unsigned lb, pb;
			memcpy(&lb, &tag.level, 4);
			memcpy(&pb, &tag.ppm, 4);
			printf("%c|%d|%lld|%lld|%u|%u|", m[i].getChannel(), m[i].getLength(), (long long)m[i].start_idx, (long long)m[i].end_idx, lb, pb);
			bool first = true;
			for (const auto &s : m[i].sentences()) {
				if (!first) putchar(' ');
				fwrite(s.data(), 1, s.size(), stdout);
				first = false;
			}
			putchar('\n');
			count++;
		}
	}
};
} // namespace

int main(int argc, char **argv) {
	if (argc < 6) {
		fprintf(stderr, "usage: %s file CU8|CF32 rate block_samples model [cpu]\n", argv[0]);
		return 2;
	}
	const Format fmt = strcmp(argv[2], "CF32") == 0 ? Format::CF32 : Format::CU8;
	const int bps = fmt == Format::CF32 ? 8 : 2;
	const int rate = atoi(argv[3]), block = atoi(argv[4]), kind = atoi(argv[5]);
	const bool cpu = argc > 6 && strcmp(argv[6], "cpu") == 0;
	FILE *f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 2; }
	std::vector<unsigned char> data;
	unsigned char buf[65536];
	size_t n;
	while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
	fclose(f);

	MemDevice dev;
	Sink sink;
	AIS::Model *model = nullptr;
	try {
		if (cpu) model = kind == 2 ? (AIS::Model *)new AIS::ModelDefault() : (kind == 0 ? (AIS::Model *)new AIS::ModelStandard() : (AIS::Model *)new AIS::ModelBase());
		else model = new AIS::ModelGPU(kind);
		for (int i = 7; i + 1 < argc; i += 2) { // -go style KEY VALUE pairs (CommandLine.cpp:196-234)
			AIS::Keys key = AIS::KEY_SETTING_PS_EMA;
			if (!strcmp(argv[i], "PS_EMA")) key = AIS::KEY_SETTING_PS_EMA;
			else if (!strcmp(argv[i], "AFC_WIDE")) key = AIS::KEY_SETTING_AFC_WIDE;
			else if (!strcmp(argv[i], "DROOP")) key = AIS::KEY_SETTING_DROOP;
			else if (!strcmp(argv[i], "FP_DS")) key = AIS::KEY_SETTING_FP_DS;
			else if (!strcmp(argv[i], "DSK")) key = AIS::KEY_SETTING_DSK;
			else if (!strcmp(argv[i], "SOXR")) key = AIS::KEY_SETTING_SOXR;
			model->SetKey(key, argv[i + 1]);
		}
		model->buildModel('A', 'B', rate, false, &dev);
	}
	catch (std::exception &e) { // what CommandLine::run does (CommandLine.cpp:752-757)
		fprintf(stderr, "config error: %s\n", e.what());
		return 4;
	}
	model->Output().out.Connect(&sink);
	const size_t step = (size_t)block * bps;
	if (getenv("ADAPTER_VARY")) { // odd, short and varying device buffers after a first one of the nominal length (network sources)
		const size_t total = data.size() / step * step; // the same samples a fixed-length run consumes
		size_t off = 0;
		int k = 0;
		while (off < total && !g_stop_requests) {
			static const int num[] = { 8, 3, 13, 1, 8, 5, 16, 2 };
			size_t n = k == 0 ? step : step * num[k & 7] / 8 + (size_t)bps * (k % 3);
			if (n > total - off) n = total - off;
			dev.push(data.data() + off, (int)n, fmt);
			off += n;
			k++;
		}
	}
	else
		for (size_t off = 0; off + step <= data.size() && !g_stop_requests; off += step) dev.push(data.data() + off, (int)step, fmt);
	fprintf(stderr, "%ld messages, %d stop requests, settings: %s\n", sink.count, g_stop_requests, model->Get().c_str());
	delete model;
	return g_stop_requests ? 3 : 0;
}
