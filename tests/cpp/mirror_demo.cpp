// Exercises the C++ mirror of the reference interface the way an SDR++ module does:
// sigpath::iqFrontEnd.init/start, sigpath::vfoManager.createVFO, reading VFO::output streams, the
// acquire/release spectrum callbacks, and a standalone dsp::channel::RxVFO::process().
//
//   mirror_demo host                       -- host-only self test of dsp::stream / dsp::block (no GPU)
//   mirror_demo run <in.cf32> <sr> <block> <fftN> <outprefix> <outSR> <bw> <off1> [<off2> ...]
//      writes <outprefix>.vfo<i>.cf32 (attached VFOs), <outprefix>.solo.cf32 (standalone RxVFO of VFO 0)
//      and <outprefix>.rows.f32 (spectrum rows, BH7 window, saturated rate sr/fftN)
//   mirror_demo pcm <in.cf32> <pcmType> <out.packet> <out.cf32>
//      SampleStreamCompressor::process on the whole file, then SampleStreamDecompressor::process on that packet
//   mirror_demo bench <sr> <block> <fftN> <nvfo> <nblocks> <readers>
//      end-to-end throughput THROUGH THE C++ INTERFACE modules use: a source thread swaps pinned cf32 blocks into the
//      input dsp::stream, sigpath::iqFrontEnd (saturated fftN-point Blackman-Harris-4 spectrum with acquire/release
//      callbacks) feeds nvfo VFOs (alternating NFM 12.5k->48k / AM 12k->24k on the bench grid) created through
//      sigpath::vfoManager, and `readers` consumer threads read()/flush() every VFO::output stream. Prints one JSON line.
#define SDRPP_SIGPATH_IMPLEMENTATION
#include <signal_path/signal_path.h>
#include <dsp/compression/sample_stream_compressor.h>
#include <dsp/compression/sample_stream_decompressor.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

static std::vector<float> g_rows;
static std::vector<float> g_rowbuf;
static float* acquireRow(void* ctx) { (void)ctx; return g_rowbuf.data(); }
static void releaseRow(void* ctx) { (void)ctx; g_rows.insert(g_rows.end(), g_rowbuf.begin(), g_rowbuf.end()); }

// a trivial Processor: proves the block/processor templates work for module-defined blocks
class Doubler : public dsp::Processor<float, float> {
    using base_type = dsp::Processor<float, float>;
public:
    Doubler(dsp::stream<float>* in) { base_type::init(in); }
    inline int process(int count, const float* in, float* o) { for (int i = 0; i < count; i++) { o[i] = 2.0f * in[i]; } return count; }
    int run() {
        int count = _in->read();
        if (count < 0) { return -1; }
        process(count, _in->readBuf, out.writeBuf);
        _in->flush();
        if (!out.swap(count)) { return -1; }
        return count;
    }
};

static int hostSelfTest() {
    dsp::stream<float> src;
    src.setBufferSize(1024);
    Doubler d(&src);
    d.out.setBufferSize(1024);
    d.start();
    d.start(); // idempotent
    double sum = 0;
    std::atomic<int> consumed{0};
    std::thread reader([&] {
        for (int b = 0; b < 50; b++) {
            int n = d.out.read();
            if (n < 0) { return; }
            for (int i = 0; i < n; i++) { sum += d.out.readBuf[i]; }
            d.out.flush();
            consumed++;
        }
    });
    for (int b = 0; b < 50; b++) {
        for (int i = 0; i < 100; i++) { src.writeBuf[i] = (float)(b + i); }
        if (b == 20) {
            // nested pause while the pipeline is idle (a block in flight during a stop is dropped, as in the reference)
            while (consumed.load() < b) { std::this_thread::yield(); }
            d.tempStop(); d.tempStop(); d.tempStart(); d.tempStart();
        }
        if (!src.swap(100)) { return 2; }
    }
    reader.join();
    d.stop();
    d.stop();
    double expect = 0;
    for (int b = 0; b < 50; b++) { for (int i = 0; i < 100; i++) { expect += 2.0 * (b + i); } }
    // stopped streams unblock: a read on a stopped stream returns -1
    src.stopReader();
    if (src.read() != -1) { return 3; }
    src.clearReadStop();
    ImGui::WaterfallVFO w;
    w.setReference(ImGui::WaterfallVFO::REF_LOWER); w.setBandwidth(2700.0); w.setOffset(1000.0);
    if (w.centerOffset != 2350.0 || w.upperOffset != 3700.0) { return 4; }
    w.setCenterOffset(0.0);
    if (w.generalOffset != -1350.0) { return 5; }
    float win[9];
    dsp::window::createWindow(dsp::window::BLACKMAN_HARRIS7, win, 8, true);
    if (!(win[4] < -0.46f && win[4] > -0.4613f)) { return 6; }
    printf("host self test ok (sum %.1f expect %.1f)\n", sum, expect);
    return sum == expect ? 0 : 1;
}

static int pcmRoundTrip(const char* inPath, int type, const char* packetPath, const char* outPath) {
    FILE* f = fopen(inPath, "rb");
    if (!f) { return 66; }
    std::vector<dsp::complex_t> x(1 << 20);
    const int count = (int)fread(x.data(), sizeof(dsp::complex_t), x.size(), f);
    fclose(f);
    if (sdrpp_cuda_init(0) != 0) { fprintf(stderr, "%s\n", sdrpp_cuda_last_error()); return 70; }
    std::vector<uint8_t> packet(8 + (size_t)count * sizeof(dsp::complex_t));
    const int bytes = dsp::compression::SampleStreamCompressor::process(count, (dsp::compression::PCMType)type, x.data(), packet.data());
    if (bytes <= 0) { fprintf(stderr, "%s\n", sdrpp_cuda_last_error()); return 70; }
    std::vector<dsp::complex_t> y((size_t)count + 1);
    dsp::compression::SampleStreamDecompressor dec;
    const int n = dec.process(bytes, packet.data(), y.data());
    f = fopen(packetPath, "wb"); fwrite(packet.data(), 1, (size_t)bytes, f); fclose(f);
    f = fopen(outPath, "wb"); fwrite(y.data(), sizeof(dsp::complex_t), (size_t)n, f); fclose(f);
    return n == count ? 0 : 1;
}

static int benchRun(double sr, int block, int fftN, int nvfo, int nblocks, int nreaders) {
    if (sdrpp_cuda_init(0) < 0) { fprintf(stderr, "%s\n", sdrpp_cuda_last_error()); return 70; }
    g_rowbuf.resize((size_t)fftN);
    static std::atomic<long long> rowsSeen{0};
    dsp::stream<dsp::complex_t> src;
    sigpath::iqFrontEnd.init(&src, sr, true, 1, false, fftN, sr / fftN, dsp::window::BLACKMAN_HARRIS4,
                             [](void*) -> float* { return g_rowbuf.data(); }, [](void*) { rowsSeen++; }, NULL);
    std::vector<VFOManager::VFO*> vfos;
    for (int i = 0; i < nvfo; i++) {
        const double off = ((double)i - (nvfo - 1) / 2.0) * (0.9 * sr / nvfo);
        const bool nfm = (i % 2) == 0;
        auto* v = sigpath::vfoManager.createVFO("vfo" + std::to_string(i), ImGui::WaterfallVFO::REF_CENTER, off, nfm ? 12500.0 : 12000.0,
                                                nfm ? 48000.0 : 24000.0, 1000.0, 200000.0, true);
        if (!v) { fprintf(stderr, "createVFO failed: %s\n", sdrpp_cuda_last_error()); return 71; }
        vfos.push_back(v);
    }
    // both halves of the input stream's double buffer hold a synthetic block (tone + pseudo-random noise); the source
    // thread then only swaps, like a driver whose DMA target is the stream buffer
    auto fill = [&](dsp::complex_t* b, uint32_t seed) {
        uint32_t s = seed;
        for (int i = 0; i < block; i++) {
            s = s * 1664525u + 1013904223u; const float a = (float)(int32_t)s * (0.01f / 2147483648.0f);
            s = s * 1664525u + 1013904223u; const float c = (float)(int32_t)s * (0.01f / 2147483648.0f);
            const float ph = 0.001f * (float)(i % 6283);
            b[i] = dsp::complex_t{ 0.3f * cosf(ph) + a, 0.3f * sinf(ph) + c };
        }
    };
    fill(src.writeBuf, 1u); fill(src.readBuf, 2u);
    std::atomic<long long> samplesOut{0};
    std::vector<std::thread> readers;
    for (int r = 0; r < nreaders; r++) {
        readers.emplace_back([&, r] {
            long long local = 0;
            while (true) {
                for (size_t i = (size_t)r; i < vfos.size(); i += (size_t)nreaders) {
                    int n = vfos[i]->output->read();
                    if (n < 0) { samplesOut += local; return; }
                    local += n;
                    vfos[i]->output->flush();
                }
            }
        });
    }
    sigpath::iqFrontEnd.start();
    const int warm = 200;   // untimed: the front end instantiates its CUDA graphs within the first ~150 blocks of a stream
    for (int b = 0; b < warm; b++) { if (!src.swap(block)) { return 74; } }
    while (sigpath::iqFrontEnd.blocksDelivered() < warm) { std::this_thread::yield(); }
    const auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < nblocks; b++) { if (!src.swap(block)) { return 74; } }
    while (sigpath::iqFrontEnd.blocksDelivered() < warm + nblocks) { std::this_thread::yield(); }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    sigpath::iqFrontEnd.stop();
    for (auto* v : vfos) { v->output->stopReader(); }
    for (auto& t : readers) { t.join(); }
    printf("{\"e2e_cpp_msps\": %.1f, \"blocks\": %d, \"block\": %d, \"vfos\": %d, \"fft\": %d, \"reader_threads\": %d, \"seconds\": %.4f, "
           "\"spectrum_rows\": %lld, \"vfo_samples_read\": %lld}\n",
           (double)nblocks * block / dt / 1e6, nblocks, block, nvfo, fftN, nreaders, dt, rowsSeen.load(), samplesOut.load());
    for (auto* v : vfos) { sigpath::vfoManager.deleteVFO(v); }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "host") { return hostSelfTest(); }
    if (argc == 8 && std::string(argv[1]) == "bench") { return benchRun(atof(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7])); }
    if (argc == 6 && std::string(argv[1]) == "pcm") { return pcmRoundTrip(argv[2], atoi(argv[3]), argv[4], argv[5]); }
    if (argc < 10 || std::string(argv[1]) != "run") { fprintf(stderr, "usage: see source\n"); return 64; }
    const std::string inPath = argv[2], prefix = argv[6];
    const double sr = atof(argv[3]);
    const int block = atoi(argv[4]), fftN = atoi(argv[5]);
    const double outSR = atof(argv[7]), bw = atof(argv[8]);
    std::vector<double> offs;
    for (int i = 9; i < argc; i++) { offs.push_back(atof(argv[i])); }

    FILE* f = fopen(inPath.c_str(), "rb");
    if (!f) { perror("input"); return 66; }
    std::vector<dsp::complex_t> x;
    { dsp::complex_t tmp[4096]; size_t n; while ((n = fread(tmp, sizeof(dsp::complex_t), 4096, f)) > 0) { x.insert(x.end(), tmp, tmp + n); } }
    fclose(f);
    const int nblocks = (int)(x.size() / (size_t)block);

    if (sdrpp_cuda_init(0) < 0) { fprintf(stderr, "%s\n", sdrpp_cuda_last_error()); return 70; }
    g_rowbuf.resize((size_t)fftN);
    dsp::stream<dsp::complex_t> src;
    sigpath::iqFrontEnd.init(&src, sr, true, 1, false, fftN, sr / fftN, dsp::window::BLACKMAN_HARRIS7, acquireRow, releaseRow, NULL);
    std::vector<VFOManager::VFO*> vfos;
    for (size_t i = 0; i < offs.size(); i++) {
        auto* v = sigpath::vfoManager.createVFO("vfo" + std::to_string(i), ImGui::WaterfallVFO::REF_CENTER, offs[i], bw, outSR, bw, bw, true);
        if (!v) { fprintf(stderr, "createVFO failed: %s\n", sdrpp_cuda_last_error()); return 71; }
        vfos.push_back(v);
    }
    if (sigpath::vfoManager.createVFO("vfo0", 1, 0, bw, outSR, bw, bw, true) != NULL) { return 72; } // duplicate name -> NULL
    if (sigpath::vfoManager.createVFO("", 1, 0, bw, outSR, bw, bw, true) != NULL) { return 73; }

    // one reader thread per VFO output stream, as a demodulator block would be
    std::vector<std::vector<dsp::complex_t>> outs(vfos.size());
    std::vector<std::thread> readers;
    for (size_t i = 0; i < vfos.size(); i++) {
        readers.emplace_back([&, i] {
            while (true) {
                int n = vfos[i]->output->read();
                if (n < 0) { return; }
                outs[i].insert(outs[i].end(), vfos[i]->output->readBuf, vfos[i]->output->readBuf + n);
                vfos[i]->output->flush();
            }
        });
    }
    sigpath::iqFrontEnd.start();
    for (int b = 0; b < nblocks; b++) {
        memcpy(src.writeBuf, x.data() + (size_t)b * block, sizeof(dsp::complex_t) * (size_t)block);
        if (b == nblocks / 2 && vfos.size() > 1) { sigpath::vfoManager.setOffset("vfo1", offs[1]); } // live retune to the same value
        if (!src.swap(block)) { return 74; }
    }
    // let the pipeline drain: one more empty-handed round trip through the input stream
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
    sigpath::iqFrontEnd.stop();
    for (auto* v : vfos) { v->output->stopReader(); }
    for (auto& t : readers) { t.join(); }
    for (size_t i = 0; i < vfos.size(); i++) {
        FILE* o = fopen((prefix + ".vfo" + std::to_string(i) + ".cf32").c_str(), "wb");
        fwrite(outs[i].data(), sizeof(dsp::complex_t), outs[i].size(), o);
        fclose(o);
    }
    { FILE* o = fopen((prefix + ".rows.f32").c_str(), "wb"); fwrite(g_rows.data(), sizeof(float), g_rows.size(), o); fclose(o); }
    for (auto* v : vfos) { sigpath::vfoManager.deleteVFO(v); }

    // standalone RxVFO::process on host pointers, in place (rx_vfo.h:89-100 passes out,out)
    {
        dsp::channel::RxVFO solo(NULL, sr, outSR, bw, offs[0]);
        std::vector<dsp::complex_t> y, work((size_t)block);
        for (int b = 0; b < nblocks; b++) {
            memcpy(work.data(), x.data() + (size_t)b * block, sizeof(dsp::complex_t) * (size_t)block);
            int n = solo.process(block, work.data(), work.data());
            if (n < 0) { fprintf(stderr, "solo: %s\n", sdrpp_cuda_last_error()); return 75; }
            y.insert(y.end(), work.begin(), work.begin() + n);
        }
        FILE* o = fopen((prefix + ".solo.cf32").c_str(), "wb");
        fwrite(y.data(), sizeof(dsp::complex_t), y.size(), o);
        fclose(o);
    }
    printf("mirror_demo ok: %d blocks, %zu VFOs, %zu spectrum rows\n", nblocks, vfos.size(), g_rows.size() / (size_t)fftN);
    return 0;
}
