// sfx_internal.h -- shared between sfx_kernels.cu (device) and sfx_abi.cu (host C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "../../include/sfx.h"

namespace sfx {

constexpr int kNfft = SFX_N_FFT;
constexpr int kHop = SFX_HOP;
constexpr int kBins = SFX_N_BINS;
constexpr int kMels = SFX_N_MELS;
constexpr int kChroma = SFX_N_CHROMA;
constexpr int kPStride = SFX_P_STRIDE;
constexpr int kTunings = SFX_N_TUNINGS;
constexpr int kWarps = 8;
constexpr int kStreamWarps = 16;         // warps per CTA of the stream kernel (sfx_stream.cu)
constexpr int kThreads = kWarps * 32;
constexpr int kExFloats = 32 * 33 * 2;   // per-warp float2[32][33] exchange tile, reused as the padded |X|^2 tile
constexpr int kP16Stride = 1056;         // halves per FP16 |X|^2 row (2112 B = 64 mod 128: conflict-free LDS.128 of the bank rows)
constexpr int kP16Row = 1024;            // halves per FP16 |X|^2 row in the scratch (bins 0..1023; the Nyquist bin travels apart): 2 KB, so that
                                         // the rows of all 296 CTAs of a 3 s batch (78.8 MB) fit the 79 MB persisting-L2 set-aside whole
constexpr int kChromaTiles = 16;         // full 8-frame tiles per pass: 32 units = 4 per warp; their K-half partial sums
                                         // (2 x 16 slots of 96 floats) fit beside the bank in the warp tiles
constexpr int kRedoCap = 128;            // peaks per clip queued for the reference-form bin (overflow is handled inline); the
                                         // list sits in the 1 KB of the warp tiles behind the key and bin arrays
constexpr int kPRow = 36;                // floats per 32 bins of the |X|^2 tile (4 pad: 128-bit conflict-free rows)
constexpr int kPartOff = 1156;           // offset of the mel partial-sum slots inside a warp's tile
constexpr int kKeyCap = 13312;           // peak keys (u32) + bins (u8) kept in shared memory during the median select
constexpr int kOrderMax = 65536;         // largest ragged batch that is processed longest-clip-first (order array in the header)
constexpr size_t kWsQueueBytes = 256;    // clip-queue counter lives in the first bytes of the workspace
constexpr size_t kWsHeader = kWsQueueBytes + 4 * static_cast<size_t>(kOrderMax);   // + the clip order of a ragged batch

struct DevTables {
    const float2* hann;     // [1024]
    const float2* tw1;      // [32][32]
    const float2* tw2;      // [32][32]
    const float2* mel_ab;   // [33][32]
    const unsigned* mel_mask;   // [32]
    const int* mel_src;     // [128][3]
    int mel_ps, mel_flush32;
    const uint16_t* chroma16;   // [100][2][12][1056] half: hi, lo * 2^11
    const float* chroma_ny;     // [100][12]
    const uint4* chroma_frag;   // [100][32 steps][2 half steps][2 hi/lo][32 lanes] A fragments of the bank
    const unsigned char* chroma_umma;   // [100][16 K blocks][4096] tcgen05 B-operand images of the bank
    const double* dctT;     // [128 mel][128 k]  (transposed: coalesced over k)
    const double* edges;    // [101]
    int sr, kmin, kmax;
};

struct Params {
    const float* wave;
    long long row_stride;
    const int* lengths;
    long long n_default;
    int B;
    int n_mfcc;
    float* out;
    long long out_stride;
    unsigned char* ws;
    long long cta_scratch_bytes;    // fused: bytes of a CTA's slice without its FP16 |X|^2 and log-mel rows; split / stream: slice / slot
    long long cta_p16_bytes;        // fused: bytes of a CTA's FP16 |X|^2 rows; all CTAs' rows form one block at ws + kWsHeader
    long long cta_lm_bytes;         // fused: bytes of a CTA's log-mel rows; one block right behind the |X|^2 block, then the slices
    int Tmax;
    long long max_samples;  // the scratch slices hold 1 + max_samples / hop frames per clip
    int aligned8;
    int max_pk;             // peak records per frame the scratch slice is sized for
    const int* order;       // clip processed q-th by the queue (ragged batches: longest first), nullptr = q
    DevTables tb;
    sfx_debug_out dbg;
};

// Samples of a clip.  A device-side length outside [1, max_samples] gives 0 and the clip gets a NaN row: its frames would
// not fit the scratch slice and its samples would be read past the row.
__device__ __forceinline__ long long clip_samples(const Params& p, int clip) {
    if (!p.lengths) return p.n_default;
    const long long n = p.lengths[clip];
    return n > p.max_samples ? 0 : n;
}

__host__ __device__ inline int rec_frames(int Tmax) { return (Tmax + kWarps - 1) / kWarps * kWarps; }

// Fused kernel workspace: header | FP16 |X|^2 rows of all CTAs | log-mel rows of all CTAs | per-CTA slices (everything else).
// The rows that phase 3 reads back are kept together so that ONE stream access-policy window can pin them in L2
// (persisting set-aside): they are written once and read once ~100 us later, and would otherwise be pushed out to HBM by
// the waveform stream in between.
// (UMMA build: the rows are stored as the shared-memory images of 128-frame x 64-bin operand blocks, 8-row atoms of 1 KB,
//  16 K blocks per atom row group: ceil(Tmax / 8) * 16 KB)
inline size_t cta_p16_bytes(int Tmax, bool umma = false) {
    if (umma) return static_cast<size_t>((Tmax + 7) / 8) * 16 * 1024;
    return (static_cast<size_t>(Tmax) * kP16Row * 2 + 255) & ~static_cast<size_t>(255);
}
inline size_t cta_lm_bytes(int Tmax) { return (static_cast<size_t>(Tmax) * kMels * 4 + 255) & ~static_cast<size_t>(255); }
// bytes of the rest of a CTA's scratch for clips of up to Tmax frames (multiple of 256)
inline size_t cta_rest_bytes(int Tmax, int max_pk) {
    // peak records (float4; one segment per warp: capacity rounded up to kWarps frames), peak keys (u32, overflow path),
    // hop energy + Nyquist + 1/scale per frame, peak bins (u8)
    size_t b = static_cast<size_t>(Tmax) * (12 + static_cast<size_t>(max_pk) * (4 + 1)) +
               static_cast<size_t>(rec_frames(Tmax)) * max_pk * 16;
    return (b + 255) & ~static_cast<size_t>(255);
}
inline size_t cta_scratch_bytes(int Tmax, int max_pk, bool umma = false) {
    return cta_p16_bytes(Tmax, umma) + cta_lm_bytes(Tmax) + cta_rest_bytes(Tmax, max_pk);
}

// ---- split pipeline (sfx_split.cu): workspace = header | per-clip peak counters | frame prefix | slices
constexpr int kSplitChunkMax = 1024;                        // clips per chunk (prep kernel = one 1024-thread block)
constexpr size_t kSplitNpkOff = 256;
constexpr size_t kSplitOffOff = kSplitNpkOff + 4 * kSplitChunkMax;
constexpr size_t kSplitHeader = kSplitOffOff + 4 * (kSplitChunkMax + 64);      // 8704 + 256 B, multiple of 256

struct SplitParams {
    Params p;              // p.ws = workspace base, p.cta_scratch_bytes = slice bytes per clip
    int chunk0;            // first clip of the chunk
    int nclips;            // clips in the chunk (<= kSplitChunkMax)
    int T_uniform;         // frames per clip when lengths == nullptr, else 0
};

// bytes of one clip's slice in the split pipeline (multiple of 256)
inline size_t split_slice_bytes(int Tmax, int max_pk) {
    size_t b = static_cast<size_t>(Tmax) * (kP16Row * 2 + kMels * 4 + 12 + 16 + static_cast<size_t>(max_pk) * (16 + 4 + 1));
    return (b + 255) & ~static_cast<size_t>(255);
}

// ---- stream pipeline (sfx_stream.cu): workspace = header (queue counter + clip order) | per CTA: stream_slots() slices
inline size_t stream_slot_bytes(int Tmax, int max_pk) {
    // per frame: FP16 |X|^2 row, log-mel row, the record of per-frame values, max_pk peak records (float4: every frame owns
    // a fixed segment, its fill level is part of the per-frame record), peak keys (u32) + bins (u8) of the overflow path
    size_t b = static_cast<size_t>(Tmax) * (kP16Row * 2 + kMels * 4 + 8 * 4 + static_cast<size_t>(max_pk) * (16 + 4 + 1));
    return (b + 255) & ~static_cast<size_t>(255);
}
size_t smem_stream();
int stream_slots();
cudaError_t configure_stream(int* blocks_per_sm);
cudaError_t launch_stream(const Params& p, int grid, bool debug, cudaStream_t stream);

size_t smem_bytes();
cudaError_t configure_split(int* frames_per_sm, int* clips_per_sm);
cudaError_t launch_split_chunk(const SplitParams& q, int grid_frames, int grid_clips, bool debug, cudaStream_t stream);
cudaError_t configure_kernels(int* blocks_per_sm);
cudaError_t launch_extract(const Params& p, int grid, bool debug, bool umma, cudaStream_t stream);
cudaError_t launch_order(const int32_t* lengths, int B, int* order, cudaStream_t stream);

// Host pipelines (sfx_extract_host*, sfx_preprocess_host_pcm16): on an early error return the pipeline's streams are drained,
// so that no copy is left in flight that still targets the caller's host buffers.
template <int N>
struct QuiesceOnError {
    cudaStream_t* streams;
    bool armed = true;
    ~QuiesceOnError() {
        if (!armed) return;
        for (int s = 0; s < N; ++s)
            if (streams[s]) cudaStreamSynchronize(streams[s]);
        cudaGetLastError();
    }
};

}  // namespace sfx
