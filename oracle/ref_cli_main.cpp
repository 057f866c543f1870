// ref_cli_main.cpp -- minimal -e / -d front end over the UNMODIFIED reference
// classes (sela::Encoder/Decoder + file::SelaFile/WavFile), i.e. what
// src/main.cpp:29-41 does minus the libao player, so that file-level byte
// parity can be checked without building libao.  TEST INFRASTRUCTURE ONLY.
#include <fstream>
#include <iostream>
#include <string>

#include "data/exception.hpp"
#include "sela/decoder.hpp"
#include "sela/encoder.hpp"

int main(int argc, char **argv)
{
    if (argc != 4) {
        std::cerr << "usage: sela_ref_cli -e in.wav out.sela | -d in.sela out.wav\n";
        return 2;
    }
    try {
        std::string mode = argv[1];
        std::ifstream in(argv[2], std::ios::binary);
        std::ofstream out(argv[3], std::ios::binary);
        if (mode == "-e") {
            file::SelaFile f = sela::Encoder(in).process();
            f.writeToFile(out);
        } else if (mode == "-d") {
            file::WavFile f = sela::Decoder(in).process();
            f.writeToFile(out);
        } else {
            return 2;
        }
    } catch (data::Exception e) {
        std::cerr << e.exceptionMessage << std::endl;
        return 1;
    }
    return 0;
}
