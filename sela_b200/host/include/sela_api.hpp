// sela_api.hpp -- the reference's operator interface for the hot path and its callers,
// re-implemented on top of the CUDA C ABI (include/sela_b200.h).  Class names, method
// names, argument meaning and error behaviour follow the reference; none of the
// computation happens on the CPU.
//
//   rice::RiceEncoder / RiceDecoder           src/include/rice.hpp:10-42
//   lpc::ResidueGenerator / SampleGenerator   src/include/lpc.hpp:86-116
//   frame::FrameEncoder / FrameDecoder        src/include/frame.hpp:8-24
//   file::WavFile / SelaFile                  src/include/file/wav_file.hpp:10-19, sela_file.hpp:10-18
//   sela::Encoder / Decoder                   src/include/sela/encoder.hpp:9-22, decoder.hpp:9-22
//   sela::Player                              src/include/sela/player.hpp:27 (stub: no audio device here)
//
// Errors from the device layer surface as `throw data::Exception(message)`, the
// reference's convention.
#pragma once

#include <fstream>

#include "sela_types.hpp"

constexpr uint8_t MAX_LPC_ORDER = 100;   // src/include/lpc.hpp:7
constexpr uint8_t CORRECTION_FACTOR = 35; // src/include/lpc.hpp:8
constexpr uint8_t MAX_RICE_PARAM = 20;   // src/include/rice.hpp:7

namespace rice {
class RiceEncoder {
    const std::vector<int32_t> &input;

public:
    explicit RiceEncoder(const data::RiceDecodedData &decodedData);
    data::RiceEncodedData process();
};
class RiceDecoder {
    const std::vector<uint32_t> &input;
    uint32_t dataCount;
    uint32_t optimumRiceParam;

public:
    explicit RiceDecoder(const data::RiceEncodedData &encodedData);
    data::RiceDecodedData process();
};
} // namespace rice

namespace lpc {
class ResidueGenerator {
    const std::vector<int32_t> &samples;
    uint8_t bitsPerSample;

public:
    explicit ResidueGenerator(const data::LpcDecodedData &data);
    data::LpcEncodedData process();
};
class SampleGenerator {
    const data::LpcEncodedData &encoded;

public:
    explicit SampleGenerator(const data::LpcEncodedData &encodedData);
    data::LpcDecodedData process();
};
} // namespace lpc

namespace frame {
class FrameEncoder {
    const data::WavFrame &wavFrame;

public:
    explicit FrameEncoder(const data::WavFrame &wavFrame);
    data::SelaFrame process();
};
class FrameDecoder {
    const data::SelaFrame &selaFrame;

public:
    explicit FrameDecoder(const data::SelaFrame &selaFrame);
    data::WavFrame process();
};
} // namespace frame

namespace file {
class WavFile {
public:
    size_t samplesPerChannelPerFrame = 2048;
    void demuxSamples();
    data::WavChunk wavChunk;
    WavFile() {}
    WavFile(uint32_t sampleRate, uint16_t bitsPerSample, uint16_t numChannels, std::vector<data::WavFrame> &&wavFrames);
    void readFromFile(std::ifstream &inputFile);
    void writeToFile(std::ofstream &outputFile);
};
class SelaFile {
public:
    data::SelaHeader selaHeader;
    std::vector<data::SelaFrame> selaFrames;
    void readFromFile(std::ifstream &inputFile);
    void writeToFile(std::ofstream &outputFile);
    SelaFile() {}
    SelaFile(uint32_t sampleRate, uint16_t bitsPerSample, uint8_t channels, std::vector<data::SelaFrame> &&selaFrames);
};
} // namespace file

namespace sela {
class Encoder {
    void readFrames();
    void processFrames(std::vector<data::SelaFrame> &encodedSelaFrames);
    std::ifstream &ifStream;
    file::WavFile wavFile;

public:
    explicit Encoder(std::ifstream &ifStream) : ifStream(ifStream) {}
    file::SelaFile process();
    // Not in the reference: process() + SelaFile::writeToFile() in one step, byte-identical output.
    // The WAV data chunk goes to the device as it lies in the file and the .sela byte stream comes
    // back ready to write (selab200_encode_container): no per-frame value structs on the host.
    void processTo(std::ofstream &outputFile);
};
class Decoder {
    void readFrames();
    void processFrames(std::vector<data::WavFrame> &decodedWavFrames);
    std::ifstream &ifStream;
    file::SelaFile selaFile;

public:
    explicit Decoder(std::ifstream &ifStream) : ifStream(ifStream) {}
    file::WavFile process();
    // Not in the reference: process() + WavFile::writeToFile() in one step, byte-identical output
    // (selab200_container_open / _decode: the .sela bytes go to the device as they lie in the file).
    void processTo(std::ofstream &outputFile);
};
// Not in the reference.  On: the processTo() drivers keep page-locked staging buffers per host thread
// and reuse them from file to file (a process that codes many files, e.g. `sela -E`); off (default):
// plain memory, released when the call returns.
void setBatchMode(bool on);

class Player {
public:
    void play(const file::WavFile &wavFile); // always throws: playback (libao) is out of scope
};
} // namespace sela
