// reftests_main.cpp -- the reference's own round-trip unit tests (test/ricetests.cpp:7-25,
// test/lpctests.cpp:10-32, test/frametests.cpp:8-70) re-expressed against this library's
// C++ mirror of the same classes, without Catch2, and with the content comparison the
// reference's frame tests lose to a moved-from vector actually performed.
// Needs a GPU.  Exit status 0 = all passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "sela_api.hpp"

#ifndef M_PI
#define M_PI 3.141592653589793238462643383279502884
#endif

static int failures = 0;
#define REQUIRE(cond)                                                      \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);  \
            failures++;                                                    \
        }                                                                  \
    } while (0)

static std::vector<int32_t> sine_deg()
{
    std::vector<int32_t> s(2048);
    for (size_t i = 0; i < s.size(); i++)
        s[i] = (int32_t)(32767 * std::sin((double)i * (M_PI / 180)));
    return s;
}

int main()
{
    try {
        { // Rice Encoder/Decoder combined test
            std::vector<int32_t> input;
            for (size_t i = 0; i < 100; i++)
                input.push_back(200 + (std::rand() % 201));
            data::RiceDecodedData in{std::vector<int32_t>(input)};
            data::RiceEncodedData enc = rice::RiceEncoder(in).process();
            data::RiceDecodedData dec = rice::RiceDecoder(enc).process();
            REQUIRE(in.decodedData.size() == dec.decodedData.size());
            REQUIRE(in.decodedData == dec.decodedData);
        }
        { // LPC Encoder/Decoder combined test
            data::LpcDecodedData in((uint8_t)16, sine_deg());
            data::LpcEncodedData enc = lpc::ResidueGenerator(in).process();
            data::LpcDecodedData dec = lpc::SampleGenerator(enc).process();
            REQUIRE(enc.optimalLpcOrder == 17); // SURVEY.md 8a KAT row 1
            REQUIRE(in.samples.size() == dec.samples.size());
            REQUIRE(in.samples == dec.samples);
        }
        { // Frame Encoder/Decoder (+ difference coding) combined test: both channels the same sine
            std::vector<std::vector<int32_t>> planes(2, sine_deg());
            data::WavFrame in((uint8_t)16, planes);
            data::SelaFrame enc = frame::FrameEncoder(in).process();
            data::WavFrame out = frame::FrameDecoder(enc).process();
            REQUIRE(enc.subFrames.size() == 2);
            REQUIRE(enc.subFrames[1].subFrameType == 1 && enc.subFrames[1].parentChannelNumber == 0);
            REQUIRE(enc.subFrames[0].residueRequiredInts == 552 && enc.subFrames[1].residueRequiredInts == 64);
            REQUIRE(in.samples.size() == out.samples.size());
            for (size_t c = 0; c < planes.size(); c++)
                REQUIRE(planes[c] == out.samples[c]);
        }
    } catch (data::Exception e) {
        std::printf("exception: %s\n", e.exceptionMessage.c_str());
        return 2;
    }
    std::printf(failures ? "%d assertion(s) failed\n" : "All tests passed%.0d\n", failures);
    return failures ? 1 : 0;
}
