// sela_types.hpp -- the value types that cross the reference's public interfaces,
// source compatible with the reference tree's src/include/data/*.hpp so that code
// written against the reference (its main.cpp, its Catch2 tests) compiles unchanged
// against this library.  Same namespaces, type names, member names and constructor
// signatures; everything else about this file is new.
//
// Reference counterparts (paths relative to the reference tree):
//   data::Exception        src/include/data/exception.hpp:7-14
//   data::WavFrame         src/include/data/wav_frame.hpp:8-17
//   data::RiceEncodedData  src/include/data/rice_encoded_data.hpp:8-21
//   data::RiceDecodedData  src/include/data/rice_decoded_data.hpp:8-16
//   data::LpcEncodedData   src/include/data/lpc_encoded_data.hpp:8-21
//   data::LpcDecodedData   src/include/data/lpc_decoded_data.hpp:8-17
//   data::SelaSubFrame     src/include/data/sela_sub_frame.hpp:7-44
//   data::SelaFrame        src/include/data/sela_frame.hpp:7-18
//   data::SelaHeader       src/include/data/sela_header.hpp:7-15
//   data::Wav*Chunk        src/include/data/wav_sub_chunk.hpp:9-33, wav_chunk.hpp:7-15
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace data {

// Thrown by value and caught by value in the CLI (src/main.cpp:101-105); deliberately
// not derived from std::exception, like the reference's.
class Exception {
public:
    const std::string exceptionMessage;
    explicit Exception(const std::string &&message) noexcept : exceptionMessage(message) {}
};

class WavFrame {
public:
    uint8_t bitsPerSample;
    std::vector<std::vector<int32_t>> samples; // [channel][sample]
    WavFrame(uint8_t bits, std::vector<std::vector<int32_t>> planes) : bitsPerSample(bits), samples(std::move(planes)) {}
};

class RiceEncodedData {
public:
    uint32_t optimumRiceParam;
    uint32_t dataCount;
    const std::vector<uint32_t> encodedData;
    RiceEncodedData(int32_t riceParam, int32_t count, const std::vector<uint32_t> &&words) noexcept
        : optimumRiceParam(riceParam), dataCount(count), encodedData(words) {}
};

class RiceDecodedData {
public:
    const std::vector<int32_t> decodedData;
    explicit RiceDecodedData(const std::vector<int32_t> &&values) noexcept : decodedData(values) {}
};

class LpcEncodedData {
public:
    uint8_t optimalLpcOrder;
    uint8_t bitsPerSample;
    const std::vector<int32_t> quantizedReflectionCoefficients;
    const std::vector<int32_t> residues;
    LpcEncodedData(uint8_t order, uint8_t bits, const std::vector<int32_t> &&q, const std::vector<int32_t> &&res) noexcept
        : optimalLpcOrder(order), bitsPerSample(bits), quantizedReflectionCoefficients(q), residues(res) {}
};

class LpcDecodedData {
public:
    uint8_t bitsPerSample;
    const std::vector<int32_t> samples;
    LpcDecodedData(uint8_t bits, const std::vector<int32_t> &&s) noexcept : bitsPerSample(bits), samples(s) {}
};

class SelaSubFrame {
public:
    uint8_t channel;
    uint8_t subFrameType;        // 0 independent, 1 difference-coded against parentChannelNumber
    uint8_t parentChannelNumber;

    uint8_t reflectionCoefficientRiceParam;
    uint16_t reflectionCoefficientRequiredInts;
    uint8_t optimumLpcOrder;
    const std::vector<uint32_t> encodedReflectionCoefficients;

    uint8_t residueRiceParam;
    uint16_t residueRequiredInts;
    uint16_t samplesPerChannel;
    const std::vector<uint32_t> encodedResidues;

    SelaSubFrame(uint8_t channel, uint8_t type, uint8_t parent, const RiceEncodedData &refl,
                 const RiceEncodedData &res) noexcept
        : channel(channel), subFrameType(type), parentChannelNumber(parent),
          reflectionCoefficientRiceParam((uint8_t)refl.optimumRiceParam),
          reflectionCoefficientRequiredInts((uint16_t)refl.encodedData.size()),
          optimumLpcOrder((uint8_t)refl.dataCount), encodedReflectionCoefficients(refl.encodedData),
          residueRiceParam((uint8_t)res.optimumRiceParam), residueRequiredInts((uint16_t)res.encodedData.size()),
          samplesPerChannel((uint16_t)res.dataCount), encodedResidues(res.encodedData) {}
};

class SelaFrame {
public:
    const int32_t syncWord = 0xAA55FF00;
    std::vector<SelaSubFrame> subFrames;
    uint8_t bitsPerSample; // not serialised
    explicit SelaFrame(uint8_t bits) : bitsPerSample(bits) {}
    uint32_t getByteCount(); // declared, never defined, in the reference as well
    void write();
};

class SelaHeader {
public:
    const uint8_t magicNumber[4] = {'S', 'e', 'L', 'a'};
    uint32_t sampleRate;
    uint16_t bitsPerSample;
    uint8_t channels;
    uint32_t numFrames;
};

class WavSubChunk {
public:
    std::string subChunkId;
    uint32_t subChunkSize;
    std::vector<int8_t> subChunkData;
};

class WavFormatSubChunk : public WavSubChunk {
public:
    int16_t audioFormat;
    uint16_t numChannels;
    uint32_t sampleRate;
    uint32_t byteRate;
    uint16_t blockAlign;
    uint16_t bitsPerSample;
};

class WavDataSubChunk : public WavSubChunk {
public:
    uint8_t bitsPerSample;
    uint8_t channels;
    std::vector<WavFrame> wavFrames;
};

class WavChunk {
public:
    std::string chunkId;
    uint32_t chunkSize;
    std::string format;
    WavFormatSubChunk formatSubChunk;
    WavDataSubChunk dataSubChunk;
    std::vector<WavSubChunk> wavSubChunks;
};

} // namespace data
