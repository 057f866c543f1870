// sela -- command-line front end over the GPU codec.  Same contract as the reference CLI
// (src/main.cpp:16-27, 53-107): `-e in.wav out.sela`, `-d in.sela out.wav`, `-p in.sela`;
// banner on stdout, data::Exception caught by value -> message on stderr, exit status 1.
// (-p needs an audio device and is not built here: it reports that and fails.)
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "sela_api.hpp"

namespace {
int usage(const std::string &prog)
{
    std::cout << "Usage: \n\n"
              << "Encoding a file:\n" << prog << " -e path/to/input.wav path/to/output.sela\n\n"
              << "Decoding a file:\n" << prog << " -d path/to/input.sela path/to/output.wav\n\n"
              << "Playing a file:\n" << prog << " -p path/to/input.sela" << std::endl;
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
        if (mode == "-e" && argc == 4) {
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
