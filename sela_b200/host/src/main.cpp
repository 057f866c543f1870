// sela -- command-line front end over the GPU codec.  Same contract as the reference CLI
// (src/main.cpp:16-27, 53-107): `-e in.wav out.sela`, `-d in.sela out.wav`, `-p in.sela`;
// banner on stdout, data::Exception caught by value -> message on stderr, exit status 1.
// (-p needs an audio device and is not built here: it reports that and fails.)
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sela_api.hpp"

namespace {
// `sela -E out_dir a.wav b.wav ...` / `sela -D out_dir a.sela b.sela ...` (not in the reference):
// many files in ONE process, so that the CUDA context is created once, staging buffers stay
// page-locked, and a few host threads read and write files while another file is on the GPU.
// Output names: <out_dir>/<input base name>.sela|.wav.  A file that fails is reported on stderr
// (reference messages) and the others still run; exit status 1 if any failed.
int run_batch(bool encode, const std::string &out_dir, const std::vector<std::string> &inputs)
{
    sela::setBatchMode(true);
    std::atomic<size_t> next{0}, failed{0};
    std::mutex log;
    unsigned workers = 8;
    if (const char *env = std::getenv("SELA_B200_WORKERS"))
        workers = (unsigned)std::max(1, std::atoi(env));
    workers = (unsigned)std::min<size_t>(workers, inputs.size());
    auto work = [&] {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= inputs.size())
                return;
            const std::string &in_path = inputs[i];
            std::string base = in_path.substr(in_path.find_last_of('/') + 1);
            const size_t dot = base.find_last_of('.');
            if (dot != std::string::npos && dot > 0)
                base.resize(dot);
            const std::string out_path = out_dir + "/" + base + (encode ? ".sela" : ".wav");
            try {
                std::ifstream in(in_path, std::ios::binary);
                if (!in)
                    throw data::Exception("cannot open input file");
                std::ofstream out(out_path, std::ios::binary);
                if (!out)
                    throw data::Exception("cannot open output file " + out_path);
                if (encode)
                    sela::Encoder(in).processTo(out);
                else
                    sela::Decoder(in).processTo(out);
            } catch (data::Exception e) {
                std::lock_guard<std::mutex> lock(log);
                std::cerr << in_path << ": " << e.exceptionMessage << std::endl;
                failed++;
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < workers; t++)
        pool.emplace_back(work);
    for (std::thread &t : pool)
        t.join();
    std::cout << (encode ? "Encoded " : "Decoded ") << inputs.size() - failed.load() << " of " << inputs.size()
              << " files" << std::endl;
    return failed.load() ? 1 : 0;
}

int usage(const std::string &prog)
{
    std::cout << "Usage: \n\n"
              << "Encoding a file:\n" << prog << " -e path/to/input.wav path/to/output.sela\n\n"
              << "Decoding a file:\n" << prog << " -d path/to/input.sela path/to/output.wav\n\n"
              << "Playing a file:\n" << prog << " -p path/to/input.sela\n\n"
              << "Many files in one process (B200 build):\n" << prog << " -E out_dir a.wav b.wav ...\n"
              << prog << " -D out_dir a.sela b.sela ..." << std::endl;
    return 0;
}
} // namespace

int main(int argc, char **argv)
{
    std::cout << "SimplE Lossless Audio v2 (B200 build). Released under MIT license" << std::endl;
    const std::string prog = argv[0];
    if (argc < 2)
        return usage(prog);
    // One GPU is all this process uses: hide the others from the driver before it starts, so that
    // cuInit does not bring up every device of an 8-GPU box (the bulk of a short run's wall time).
    if (!std::getenv("CUDA_VISIBLE_DEVICES")) {
        const char *dev = std::getenv("SELAB200_DEVICE");
        setenv("CUDA_VISIBLE_DEVICES", dev ? dev : "0", 1);
        setenv("SELAB200_DEVICE", "0", 1);
    }
    // SELA_B200_CLASSIC=1: the reference's two-step call sequence (process(), then writeToFile())
    // instead of the fused file-to-file drivers; same bytes, more host work.
    const bool classic = std::getenv("SELA_B200_CLASSIC") != nullptr;
    try {
        const std::string mode = argv[1];
        if ((mode == "-E" || mode == "-D") && argc >= 4) {
            std::cout << (mode == "-E" ? "Encoding " : "Decoding ") << argc - 3 << " files into " << argv[2] << std::endl;
            const int rc = run_batch(mode == "-E", argv[2], std::vector<std::string>(argv + 3, argv + argc));
            std::cout.flush();
            std::cerr.flush();
            std::_Exit(rc);
        } else if (mode == "-e" && argc == 4) {
            std::ifstream in(argv[2], std::ios::binary);
            std::ofstream out(argv[3], std::ios::binary);
            std::cout << "Encoding: " << argv[2] << std::endl;
            if (classic) {
                file::SelaFile coded = sela::Encoder(in).process();
                coded.writeToFile(out);
            } else {
                sela::Encoder(in).processTo(out);
            }
        } else if (mode == "-d" && argc == 4) {
            std::ifstream in(argv[2], std::ios::binary);
            std::ofstream out(argv[3], std::ios::binary);
            std::cout << "Decoding: " << argv[2] << std::endl;
            if (classic) {
                file::WavFile pcm = sela::Decoder(in).process();
                pcm.writeToFile(out);
            } else {
                sela::Decoder(in).processTo(out);
            }
        } else if (mode == "-p" && argc == 3) {
            std::ifstream in(argv[2], std::ios::binary);
            std::cout << "Playing: " << argv[2] << std::endl;
            file::WavFile pcm = sela::Decoder(in).process();
            sela::Player().play(pcm);
        } else {
            return usage(prog);
        }
    } catch (data::Exception e) {
        std::cerr << e.exceptionMessage << std::endl;
        return 1;
    }
    // Output files are closed (their streams went out of scope above).  Leave without tearing the
    // CUDA context down piece by piece: the driver reclaims everything with the process, and the
    // orderly teardown of a context holding ~1 GB costs a few hundred ms.
    std::cout.flush();
    std::cerr.flush();
    std::_Exit(0);
}
