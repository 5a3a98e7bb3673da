// kernels.cuh -- launch-side declarations of the sm_100a kernels (definitions in kernels.cu).
#ifndef KPEG_KERNELS_CUH
#define KPEG_KERNELS_CUH

#include <cuda_runtime.h>

#include "kpeg_common.h"

namespace kpeg {

constexpr int MAX_RELAY_ROUNDS = 64; // slots in DevMeta::changed

// DevMeta::changed slot of relay round r (r >= 1).  Rounds beyond the array reuse slots 2..63 (the
// host zeroes a slot before reusing it); 62 is even, so slot parity == round parity, which is what
// selects the ping-pong work list.
inline int relay_slot(int r) { return r < MAX_RELAY_ROUNDS ? r : 2 + ((r - 2) % (MAX_RELAY_ROUNDS - 2)); }

// Device-resident bookkeeping of one job; zeroed before every decode, read back once at the end.
struct DevMeta {
    uint32_t total_kept;   // unstuffed bytes
    uint32_t total_rst;    // RSTn markers found
    uint32_t total_bits;   // 8 * total_kept
    uint32_t nsub;         // subsequences actually used
    uint32_t status;       // ST_* bits
    uint32_t exact_samples; // tie records written to the strips' own slots of the list (statistics)
    uint32_t colour_exact;
    uint32_t final_slot;   // absolute slot the last subsequence ended on
    uint32_t tie_records;  // records in the shared tail of the tie list (beyond the strips' own slots)
    uint32_t tie_inline;   // pixels resolved inside K3 because the record buffer was full
    uint32_t relay_rounds; // last relay round that ran (device-side round loop)
    uint32_t grid_bar;     // arrival counter of the relay loop's grid barrier
    uint32_t rec_alt_count; // private record areas handed out by the sparse relay rounds
    uint32_t changed[MAX_RELAY_ROUNDS];
};

struct UnstuffArgs {
    const uint8_t *scan; // stuffed entropy-coded bytes (device)
    uint32_t scan_len;
    uint32_t ntiles;
    uint32_t *tile_kept, *tile_rst;      // per-tile counts, then (after the scan kernel) exclusive offsets
    uint32_t *cls;                       // [ntiles * UNSTUFF_THREADS] per 16-byte chunk: keep mask | RSTn mask << 16
    uint8_t *words;                      // unstuffed stream as big-endian 32-bit words (byte address ^ 3)
    uint32_t *seg_bit;                   // [nseg + 2]
    uint32_t nseg;                       // expected number of segments
    DevMeta *meta;
};

constexpr int UNSTUFF_THREADS = 256;
constexpr int UNSTUFF_BYTES_PER_THREAD = 16;
constexpr int UNSTUFF_TILE = UNSTUFF_THREADS * UNSTUFF_BYTES_PER_THREAD;

struct EntropyArgs {
    const uint32_t *words;
    const uint32_t *seg_bit;
    const DeviceTables *tables;
    DevMeta *meta;
    SubState *state;      // [nsub_max] relay states X[i]
    uint32_t *worklist[2]; // [nsub_max] each: subsequences whose input changed in the previous relay round
    uint32_t *seg_hint;   // [nsub_max]
    uint32_t *start_slot; // [nsub_max] absolute slot at the entry of subsequence i
    uint2 *scan_tiles;    // [ceil(nsub_max / 1024)] per-tile aggregates / carries of the offset scan
    uint32_t *rec;        // symbol records, [group of 32 subsequences][rec_kmax][32] (nullptr = records off)
    uint32_t *nrec;       // [nsub_max] records per subsequence | (private area index + 1) << 10
    uint32_t *rec_alt;    // [rec_alt_cap][rec_kmax] private record areas of subsequences redone in sparse rounds
    uint32_t rec_alt_cap;
    uint32_t rec_kmax;
    int16_t *coef;        // [total_blocks][64]
    int16_t *dcdiff;      // [total_blocks]
    uint32_t nsub_max;
    JobGeom g;
};

constexpr int ENTROPY_THREADS = 128;

struct DcArgs {
    const int16_t *dcdiff;
    int16_t *dc;
    int32_t *tile_carry; // [ntiles][4]: per component running value at the end of the tile, [3] = "tile contains a reset"
    uint32_t ntiles;
    JobGeom g;
};
constexpr int DC_THREADS = 256;
constexpr int DC_MCUS_PER_THREAD = 4;
constexpr int DC_TILE = DC_THREADS * DC_MCUS_PER_THREAD;

struct IdctArgs {
    const int16_t *coef;
    const int16_t *dc;
    const int16_t *dcdiff;
    const DeviceTables *tables;
    uint8_t *pixels; // [nimages][height][width][ncomp]
    DevMeta *meta;
    uint4 *tie_rec;       // [tie_cap] pixels with at least one sample inside the tie band: 16 slots per strip, then a shared tail
    uint32_t tie_cap;
    uint32_t *overflow_mcu; // [strips] set when a strip had tied pixels that did not fit tie_rec
    JobGeom g;
};
constexpr int IDCT_MCUS_PER_CTA = 32;

void kernels_context_created();   // contexts of this process share the device: the cooperative relay loops of
void kernels_context_destroyed(); // all of them together must stay co-resident
void kernels_configure(int max_concurrent_jobs); // jobs (lanes) that may be on the device at the same time // per-device function attributes (call once after cudaSetDevice)
void launch_unstuff(const UnstuffArgs &a, uint32_t sub_bits, cudaStream_t s, uint32_t *launches);
void launch_entropy_cold(const EntropyArgs &a, cudaStream_t s, uint32_t *launches);
void launch_entropy_relay(const EntropyArgs &a, int round, cudaStream_t s, uint32_t *launches);
// rounds first..last (first >= 2) in ONE cooperative launch: a device-side loop with a grid barrier
// between rounds, stopping early at the fixed point; DevMeta::relay_rounds = last round that ran
// fails (nothing launched) when the driver refuses the cooperative launch; the caller then issues the rounds one by one
cudaError_t launch_entropy_relay_loop(const EntropyArgs &a, int first, int last, cudaStream_t s, uint32_t *launches);
void launch_entropy_scan(const EntropyArgs &a, cudaStream_t s, uint32_t *launches);
void launch_entropy_write(const EntropyArgs &a, cudaStream_t s, uint32_t *launches);  // Huffman final pass
void launch_entropy_expand(const EntropyArgs &a, cudaStream_t s, uint32_t *launches); // record final pass
void launch_dc_scan(const DcArgs &a, cudaStream_t s, uint32_t *launches);
cudaError_t launch_idct(const IdctArgs &a, cudaStream_t s, uint32_t *launches); // fails without a tensor-map encoder in the driver
void launch_idct_patch(const IdctArgs &a, cudaStream_t s, uint32_t *launches);
void launch_merge_dc(int16_t *coef_out, const int16_t *coef, const int16_t *dc, const int16_t *dcdiff, uint32_t nblocks,
                     uint32_t flags, cudaStream_t s);

} // namespace kpeg
#endif
