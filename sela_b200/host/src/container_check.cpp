// container_check -- CPU-only exercise of the container code (no GPU needed):
//   container_check wav  in.wav  out.wav    file::WavFile::readFromFile -> writeToFile
//   container_check sela in.sela out.sela   file::SelaFile::readFromFile -> writeToFile
// Used by tests/test_host_logic.py to compare against the reference's own reader/writer.
#include <fstream>
#include <iostream>
#include <string>

#include "sela_api.hpp"

int main(int argc, char **argv)
{
    if (argc != 4)
        return 2;
    try {
        std::ifstream in(argv[2], std::ios::binary);
        std::ofstream out(argv[3], std::ios::binary);
        const std::string mode = argv[1];
        if (mode == "wav") {
            file::WavFile f;
            f.readFromFile(in);
            const data::WavFormatSubChunk &fmt = f.wavChunk.formatSubChunk;
            // re-wrap the demuxed frames exactly as the decoder would (canonical header, tail dropped)
            file::WavFile canon(fmt.sampleRate, fmt.bitsPerSample, fmt.numChannels,
                                std::move(f.wavChunk.dataSubChunk.wavFrames));
            canon.writeToFile(out);
        } else if (mode == "sela") {
            file::SelaFile f;
            f.readFromFile(in);
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
