// sela -- command-line front end over the GPU codec.  Same contract as the reference CLI
// (src/main.cpp:16-27, 53-107): `-e in.wav out.sela`, `-d in.sela out.wav`, `-p in.sela`;
// banner on stdout, data::Exception caught by value -> message on stderr, exit status 1.
// (-p needs an audio device and is not built here: it reports that and fails.)
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
    try {
        const std::string mode = argv[1];
        if (mode == "-e" && argc == 4) {
            std::ifstream in(argv[2], std::ios::binary);
            std::ofstream out(argv[3], std::ios::binary);
            std::cout << "Encoding: " << argv[2] << std::endl;
            file::SelaFile coded = sela::Encoder(in).process();
            coded.writeToFile(out);
        } else if (mode == "-d" && argc == 4) {
            std::ifstream in(argv[2], std::ios::binary);
            std::ofstream out(argv[3], std::ios::binary);
            std::cout << "Decoding: " << argv[2] << std::endl;
            file::WavFile pcm = sela::Decoder(in).process();
            pcm.writeToFile(out);
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
    return 0;
}
