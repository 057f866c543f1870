/*
 * sela_b200.h -- C ABI of the B200-native SELA per-frame encode/decode hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI; its
 * boundary is a set of C++ classes over std::vector-owning value structs.  Each
 * entry point below names the reference interface it replaces (paths relative
 * to the reference tree) and is what a reference-side binding would call; the
 * C++ mirror of the reference classes that sits on top of it lives in
 * sela_b200/host/ (see INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer it passes;
 *     the library owns device memory, streams and its workspace.
 *   - every function returns SELAB200_OK (0) or a negative selab200_status;
 *     selab200_last_error() returns a thread-local description.  The C++ shim
 *     turns a non-zero status into `throw data::Exception(...)`, the reference's
 *     error convention (src/include/data/exception.hpp:7-14, src/main.cpp:101-105).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with SELAB200_ERR_NO_DEVICE.
 *   - one process drives one GPU (selab200_init(device)) or several
 *     (selab200_init_devices); calls are synchronous unless they take a stream
 *     (the *_device forms), and are serialised by an internal mutex -- inside a
 *     call every device works on its own thread with its own context.
 *   - a "subframe" is one channel of one 2048-sample frame
 *     (src/include/file/wav_file.hpp:12); frames are independent.
 */
#ifndef SELA_B200_H_
#define SELA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELAB200_ABI_VERSION     2 /* 2: selab200_init_devices, selab200_device_count, selab200_rice_decode_flagged */
#define SELAB200_FRAME_SAMPLES   2048 /* src/include/file/wav_file.hpp:12 */
#define SELAB200_MAX_LPC_ORDER   100  /* src/include/lpc.hpp:7            */
#define SELAB200_MAX_RICE_PARAM  20   /* src/include/rice.hpp:7           */
#define SELAB200_MAX_CHANNELS    16

typedef enum selab200_status {
    SELAB200_OK              = 0,
    SELAB200_ERR_NO_DEVICE   = -1, /* no CUDA device / extension not usable        */
    SELAB200_ERR_CUDA        = -2, /* a CUDA runtime call failed                    */
    SELAB200_ERR_ARGUMENT    = -3, /* null pointer, bad channel count, ...          */
    SELAB200_ERR_CAPACITY    = -4, /* output word arena too small                   */
    SELAB200_ERR_RANGE       = -5, /* sample outside the 16-bit domain              */
    SELAB200_ERR_BITSTREAM   = -6, /* malformed subframe descriptor / Rice stream   */
    SELAB200_ERR_NOT_INIT    = -7
} selab200_status;

/* One coded subframe.  Field-for-field data::SelaSubFrame
 * (src/include/data/sela_sub_frame.hpp:7-29) with the two word vectors replaced
 * by offsets (in uint32 words) into a flat arena.  32 bytes, naturally aligned. */
typedef struct selab200_subframe_desc {
    uint8_t  channel;
    uint8_t  subframe_type;     /* 0 independent, 1 difference-coded            */
    uint8_t  parent_channel;
    uint8_t  refl_rice_param;   /* reflectionCoefficientRiceParam               */
    uint16_t refl_words;        /* reflectionCoefficientRequiredInts            */
    uint8_t  lpc_order;         /* optimumLpcOrder                              */
    uint8_t  res_rice_param;    /* residueRiceParam                             */
    uint16_t res_words;         /* residueRequiredInts                          */
    uint16_t samples;           /* samplesPerChannel                            */
    uint32_t reserved;          /* zero                                         */
    uint64_t refl_offset;       /* encodedReflectionCoefficients -> words[...]  */
    uint64_t res_offset;        /* encodedResidues               -> words[...]  */
} selab200_subframe_desc;

/* ------------------------------------------------------------ life cycle -- */

/* Bind this process to CUDA device `device` (>= 0) and create the streams and
 * workspace.  Idempotent for the same device; a different device than before
 * tears the old context down first (its streams and pools belong to the old device). */
int  selab200_init(int device);
/* Several devices of one box (selagpu_init(device_count, device_ids)): every device gets a
 * context of its own (streams, events, pools).  The host-buffer batch calls (selab200_encode_frames,
 * selab200_decode_frames, selab200_encode_container, selab200_container_decode) then cut the frames into one contiguous block per device
 * -- n/D frames each, the last device takes the rest, the way sela::Encoder::processFrames cuts them for its
 * threads (src/sela/encoder.cpp:58-73) -- and run the blocks concurrently, reading and writing disjoint ranges
 * of the caller's buffers; results are byte-identical to a single device's.  devices[0] is the primary: it
 * serves the stage-level calls and holds the byte image of an open container.  The *_device forms run on whichever initialised device
 * owns the buffers they are given. */
int  selab200_init_devices(int count, const int *devices);
int  selab200_device_count(void);
void selab200_shutdown(void);
const char *selab200_last_error(void);
int  selab200_abi_version(void);
/* Number of kernel launches issued by this process so far (bench bookkeeping). */
uint64_t selab200_launch_count(void);

/* Device self-test: counts inputs s in [-65535, 65535] for which the kernels'
 * division-free s/32767 differs from IEEE division (must be 0). */
int selab200_selftest(uint32_t *mismatches);

/* Pinned host memory for the host-buffer calls (pageable memory also works, slower). */
void *selab200_host_alloc(size_t bytes);
void  selab200_host_free(void *p);

/* Safe arena size (in words) for encoding n_frames x channels subframes of 16-bit
 * audio: every stream the Rice coder can produce for in-domain input fits. */
size_t selab200_encode_words_bound(uint32_t n_frames, uint32_t channels);

/* ------------------------------------------------- batch coder (frames) -- */

/* Replaces sela::Encoder::processFrames (src/sela/encoder.cpp:40-92), i.e.
 * frame::FrameEncoder::process (src/frame/frame_encoder.cpp:11-102) over every
 * frame.  pcm: interleaved little-endian int16 exactly as the WAV data chunk
 * (src/file/wav_file.cpp:194-200), n_frames*2048*channels samples.  channels==2
 * triggers the difference-coding decision for channel 1.  descs:
 * n_frames*channels entries in frame order, channel order.  words: arena of
 * words_capacity uint32; subframes are laid out in file order (refl words then
 * residue words, src/file/sela_file.cpp:120-135); *words_used receives the total. */
int selab200_encode_frames(const int16_t *pcm, uint32_t n_frames, uint32_t channels,
                           selab200_subframe_desc *descs, uint32_t *words,
                           size_t words_capacity, size_t *words_used);

/* Replaces sela::Decoder::processFrames (src/sela/decoder.cpp:41-92), i.e.
 * frame::FrameDecoder::process (src/frame/frame_decoder.cpp:11-72) over every
 * frame.  pcm_out: interleaved int16, n_frames*2048*channels samples (the
 * layout file::WavFile::writeToFile emits, src/file/wav_file.cpp:244-266).
 * Descriptors are validated (order <= 100, rice params < 32, samples == 2048,
 * channel/parent < channels, offsets inside n_words); invalid input returns
 * SELAB200_ERR_BITSTREAM instead of the reference's undefined behaviour. */
int selab200_decode_frames(const selab200_subframe_desc *descs, uint32_t n_frames,
                           uint32_t channels, const uint32_t *words, size_t n_words,
                           int16_t *pcm_out);

/* Device-resident forms: every pointer is a device pointer, work is enqueued on
 * `stream` (a cudaStream_t; NULL is the legacy default stream, as everywhere in CUDA) and the call
 * returns without synchronising.  d_status (int32, device) receives 0 or a
 * selab200_status once the stream has drained; d_words_used is a device uint64.
 * workspace: selab200_*_workspace_bytes() bytes of device memory, 256-aligned.  d_words must be
 * 16-byte aligned and readable up to the next 16-byte boundary past its last word (the Rice
 * decoder fetches 16 bytes at a time); cudaMalloc / torch allocations satisfy both.  d_pcm (stereo) must
 * be 16-byte aligned as well. */
size_t selab200_encode_workspace_bytes(uint32_t n_frames, uint32_t channels);
int selab200_encode_frames_device(const int16_t *d_pcm, uint32_t n_frames, uint32_t channels,
                                  selab200_subframe_desc *d_descs, uint32_t *d_words,
                                  size_t words_capacity, uint64_t *d_words_used,
                                  int32_t *d_status, void *d_workspace, size_t workspace_bytes,
                                  void *stream);
size_t selab200_decode_workspace_bytes(uint32_t n_frames, uint32_t channels);
int selab200_decode_frames_device(const selab200_subframe_desc *d_descs, uint32_t n_frames,
                                  uint32_t channels, const uint32_t *d_words, size_t n_words,
                                  int16_t *d_pcm_out, int32_t *d_status, void *d_workspace,
                                  size_t workspace_bytes, void *stream);

/* The Rice-decode kernel (K5 of SURVEY.md 2) on its own, device resident: the residue streams of
 * every subframe -> d_residues[subframe][2048] int32, i.e. rice::RiceDecoder::process
 * (src/rice/rice_decoder.cpp:54-61) over a batch.  Algorithmic bytes: the words read + 4 B per
 * sample written.  Asynchronous on `stream`. */
int selab200_rice_decode_frames_device(const selab200_subframe_desc *d_descs, uint32_t n_frames,
                                       uint32_t channels, const uint32_t *d_words, size_t n_words,
                                       int32_t *d_residues, int32_t *d_status, void *stream);
/* How many streams of the last selab200_rice_decode_frames_device call the fast decoder handed to the
 * general lane-per-stream parser (streams it could not split, or that hold a symbol its windows do not
 * cover; results are identical either way).  Synchronises.  Diagnostics / bench bookkeeping. */
int selab200_rice_decode_flagged(uint32_t *n_flagged);

/* ------------------------------------------------ .sela container level -- */

/* The byte-packed file format of file::SelaFile (src/file/sela_file.cpp:19-137): 15-byte header
 * ("SeLa", sampleRate u32, bitsPerSample u16, channels u8, numFrames u32, little endian), then
 * per frame the sync word 0xAA55FF00 and per subframe
 *   channel, type, parent, reflK (u8), reflInts (u16), order (u8), refl words,
 *   resK (u8), resInts (u16), samples (u16), residue words.
 * These entry points move whole containers: the encoder's gather kernel writes the byte stream
 * in place (headers included), the decoder realigns the word arrays on the device, so the host
 * never builds per-subframe vectors.  Bytes are identical to what
 * file::SelaFile::writeToFile emits for the frames selab200_encode_frames returns. */
typedef struct selab200_container_info {
    uint32_t sample_rate;
    uint16_t bits_per_sample;
    uint8_t  channels;
    uint8_t  reserved;
    uint32_t header_frames;  /* the numFrames field                                            */
    uint32_t n_frames;       /* frames present: the walk stops quietly at the first bad sync
                                word, as the reference reader does (sela_file.cpp:48-56)       */
    uint64_t n_words;        /* Rice words of those frames                                     */
    uint64_t n_bytes_used;   /* bytes of the container those frames (and the header) occupy    */
} selab200_container_info;

/* Worst-case container size for n_frames x channels subframes of 16-bit audio. */
size_t selab200_container_bound(uint32_t n_frames, uint32_t channels);

/* WAV data chunk -> complete .sela byte stream: sela::Encoder::process + file::SelaFile::writeToFile
 * (src/sela/encoder.cpp:94-99, src/file/sela_file.cpp:105-137).  pcm as in selab200_encode_frames.
 * SELAB200_ERR_CAPACITY if `capacity` bytes do not suffice (*bytes_used = the size needed). */
int selab200_encode_container(const int16_t *pcm, uint32_t n_frames, uint32_t channels,
                              uint32_t sample_rate, uint16_t bits_per_sample,
                              uint8_t *container, size_t capacity, size_t *bytes_used);

/* Host-only parse of a container (no device needed): header fields and the frame walk of
 * file::SelaFile::readFromFile (src/file/sela_file.cpp:19-103).  Errors (SELAB200_ERR_BITSTREAM)
 * carry the message the host mirror throws for the same file: too small, bad magic, truncated. */
int selab200_container_info_get(const uint8_t *container, size_t n_bytes, selab200_container_info *info);

/* Same walk, also returning where each frame starts: offsets[i] = byte position of frame i's sync
 * word, offsets[n_frames] = one past the last frame.  Frames are independent, so a byte range
 * [offsets[a], offsets[b]) behind a 15-byte header that says b-a frames is itself a container --
 * which is how a file is cut into per-GPU blocks (sela_b200/distributed.py).  Host only.
 * SELAB200_ERR_CAPACITY if capacity < n_frames + 1 (info is filled in either way). */
int selab200_container_frame_offsets(const uint8_t *container, size_t n_bytes, uint64_t *offsets,
                                     size_t capacity, selab200_container_info *info);

/* Complete .sela byte stream -> interleaved int16 PCM (the WAV data chunk):
 * file::SelaFile::readFromFile + sela::Decoder::processFrames.  open: parses the header, starts
 * the upload and walks the frame headers on the host meanwhile (the walk is a pointer chase
 * through the byte stream -- inherently serial, microseconds per thousand frames; everything
 * that touches the payload runs on the device).  `container` must stay valid until close.
 * decode: pcm_out receives info.n_frames * channels * 2048 samples. */
typedef struct selab200_container selab200_container;
int  selab200_container_open(const uint8_t *container, size_t n_bytes, selab200_container **handle,
                             selab200_container_info *info);
int  selab200_container_decode(selab200_container *handle, int16_t *pcm_out);
void selab200_container_close(selab200_container *handle);

/* ------------------------------------------ stage level (host buffers) -- */

/* lpc::ResidueGenerator::process (src/lpc/residue_generator.cpp:121-134) for
 * n_sub independent 2048-sample signals.  samples: [n_sub][2048] int32 in the
 * 17-bit domain |s| <= 65535 (16-bit channels and their L-R difference).
 * Outputs: order[n_sub]; q[n_sub][100] (first order[i] valid, rest 0);
 * residues[n_sub][2048]. */
int selab200_lpc_residues(const int32_t *samples, uint32_t n_sub, uint8_t *order,
                          int32_t *q, int32_t *residues);

/* lpc::SampleGenerator::process (src/lpc/sample_generator.cpp:32-39). */
int selab200_lpc_samples(const int32_t *residues, uint32_t n_sub, const uint8_t *order,
                         const int32_t *q, int32_t *samples);

/* rice::RiceEncoder::process (src/rice/rice_encoder.cpp:73-81) for n_streams
 * independent inputs.  values: [n_streams][stride] int32, counts[i] <= stride
 * <= 2048 used from row i.  Outputs per stream: rice_param, n_words, and the
 * words at words[i*words_stride ...]; SELAB200_ERR_CAPACITY if a stream needs
 * more than words_stride words (n_words[] still holds the required sizes). */
int selab200_rice_encode(const int32_t *values, const uint32_t *counts, uint32_t n_streams,
                         uint32_t stride, uint32_t *rice_param, uint32_t *n_words,
                         uint32_t *words, uint32_t words_stride);

/* rice::RiceDecoder::process (src/rice/rice_decoder.cpp:54-61) for n_streams
 * inputs: words[i*words_stride ... +n_words[i]) -> out[i*out_stride ... +counts[i]).
 * Reads past n_words[i] see zero bits (the reference would read out of bounds). */
int selab200_rice_decode(const uint32_t *words, const uint32_t *n_words, uint32_t words_stride,
                         const uint32_t *rice_param, const uint32_t *counts, uint32_t n_streams,
                         int32_t *out, uint32_t out_stride);

#ifdef __cplusplus
}
#endif
#endif /* SELA_B200_H_ */
