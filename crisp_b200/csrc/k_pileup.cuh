// k_pileup.cuh — the pileup of the reference's front end on the device: per-position, per-pool,
// per-strand allele counts from BAM alignments, the candidate screen, and the packing of the candidate sites into the
// engine's batch arrays — everything calculate_allelecounts (parsebam/variant.c:225-342) + advance_read (:56-131) + the
// mismatch count of parse_cigar (parsebam/bamread.c:10-76) + sort_allelecounts / the potential-variant test
// (variantcalls.c:62-92) + the three read-list walks of host/crisp_shim.c produce for a window of reference positions, so
// that the alignments (in the BAM's own encoding) are the only thing that crosses PCIe.
//
// Scope of this version (checked on the device by k_pile_check, never assumed): substitutions only — alignments whose CIGAR
// is one match block between clips, no overlapping mate pairs (the reference's OVERLAPPING_PE_READS merge, variant.c:137-220,
// and its indel allele discovery stay on the host).
//
// What the reference does per (position p, pool j, read) and what is reproduced here bit for bit:
//   a read is visible at p when read.position <= p < read.lastpos            (variant.c:251-257: lastpos is one past the
//                                                                              last aligned base; advance_read fails at it)
//   readdepths[j]++ and MQcounts[class of the mapping quality]++ for every visible read          (:280-283)
//   MAX_MM = alignedbases/40 + 2; mismatches = bases of the match block that differ from the reference (either case) and
//   are not 'N'                                                                (:286, bamread.c:31-41)
//   filteredreads[0]++ when mismatches > MAX_MM and mquality >= MIN_M, else [1]++ when quality < MINQ and mquality >= MIN_M (:288-289)
//   the base is counted when quality >= MINQ and mquality >= MIN_M and mismatches <= MAX_MM and
//   (base != refb or mismatches <= MAX_MM - 1): indcounts[j][BTI[base] + 8*strand]++             (:299-312)
//   counted reads are the likelihood stream (filter '0', type 0) and, per tested allele, the alt-read list in read order.
#pragma once
#include "dmath.cuh"
#include "engine_types.cuh"

namespace crisp {

struct PileReads {   // alignments of all pools, pool-major, every pool in BAM (coordinate) order
	int P;
	uint32_t n_reads;
	unsigned long long n_bases;
	const uint32_t* pool_off;      // [P+1]
	const crispgpu_alnrec* rec;    // [R]
	const int32_t* mate_pos;       // [R] or null
	const int32_t* isize;          // [R] or null
	const uint8_t* seq4;           // 4-bit bases, read r at seq4 + base_off/2
	const uint8_t* qual;           // phred scores, read r at qual + base_off
	uint8_t* flags;                // [R] written by k_pile_tile: bit 0 passes the read filters, bit 1 also when the base equals the reference
	int32_t* state;                // [0] error bits of k_pile_check, [1] longest match block, [2] candidates
};

enum { kPileErrAlign = 1, kPileErrBounds = 2, kPileErrOrder = 4, kPileErrMate = 8, kPileErrPool = 16 };

struct PileWindow {
	int tid, w0, W;            // reference positions [w0, w0 + W)
	int ref0;                  // reference position of refseq[0]
	int reflen;
	const uint8_t* refseq;     // reference characters (either case) covering the window and the reads that overlap it
	int min_q, min_m;          // MINQ, MIN_M
};

// BAM's 4-bit base code -> the character bam_nt16_rev_table gives the reference's reader (bamread.c:122) and BTI of it
// (variant.c:17-26): "=ACMGRSVTWYHKDBN" -> 0 0 1 0 2 0 1 0 3 0 1 0 2 0 1 0
__device__ __forceinline__ int nt16_char(int c) { return (int)(((c & 8) ? 0x4e42444b48595754ull : 0x565352474d43413dull) >> (8 * (c & 7))) & 255; }
__device__ __forceinline__ int nt16_bti(int c) { return (int)(0x0102010301020100ull >> (4 * c)) & 3; }
__device__ __forceinline__ int upper(int ch) { return (ch >= 'a' && ch <= 'z') ? ch - 32 : ch; }
__device__ __forceinline__ int ref_code(int up) { return up == 'A' ? 0 : up == 'C' ? 1 : up == 'G' ? 2 : up == 'T' ? 3 : 255; }

__device__ __forceinline__ int ref_at(const PileWindow& w, int p)
{
	const int k = p - w.ref0;
	return (k >= 0 && k < w.reflen) ? upper(__ldg(w.refseq + k)) : 'N';
}

__device__ __forceinline__ uint32_t lower_bound_pos(const crispgpu_alnrec* rec, uint32_t lo, uint32_t hi, int key)
{
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (rec[mid].pos < key) lo = mid + 1; else hi = mid; }
	return lo;
}

// every alignment against the scope and the layout rules of the interface; the longest match block of the set
__global__ void k_pile_check(PileReads rd)
{
	int err = 0, span = 0;
	for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < rd.n_reads; r += gridDim.x * blockDim.x) {
		const crispgpu_alnrec a = rd.rec[r];
		if (a.base_off & 7u) err |= kPileErrAlign;
		if ((unsigned long long)a.base_off + (((unsigned)a.clip + a.mlen + 7u) & ~7u) > rd.n_bases || a.mlen == 0) err |= kPileErrBounds;
		if (r > 0 && rd.rec[r - 1].pos > a.pos) {
			// allowed only across a pool boundary
			bool boundary = false;
			uint32_t lo = 0, hi = rd.P;
			while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (rd.pool_off[mid] < r) lo = mid + 1; else hi = mid; }
			boundary = lo <= (uint32_t)rd.P && rd.pool_off[lo] == r;
			if (!boundary) err |= kPileErrOrder;
		}
		if (rd.mate_pos) {
			const int mp = rd.mate_pos[r], is = rd.isize ? rd.isize[r] : 0;
			// the mate starts inside this read's match block: the reference would merge the pair (variant.c:274)
			if (mp > a.pos && mp < a.pos + (int)a.mlen - 1) err |= kPileErrMate;
			// the reference's "2nd read ends before 1st read" filter could fire (variant.c:271)
			if (is < 0 && a.pos < mp) err |= kPileErrMate;
		}
		span = max(span, (int)a.mlen);
	}
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		if (rd.pool_off[0] != 0 || rd.pool_off[rd.P] != rd.n_reads) err |= kPileErrPool;
		for (int j = 0; j < rd.P; j++) if (rd.pool_off[j] > rd.pool_off[j + 1]) err |= kPileErrPool;
	}
	err = __reduce_or_sync(0xffffffffu, err);
	span = __reduce_max_sync(0xffffffffu, span);
	if ((threadIdx.x & 31) == 0) {
		if (err) atomicOr(rd.state, err);
		atomicMax(rd.state + 1, span);
	}
}

constexpr int kPileTile = 1024;   // reference positions per CTA tile
constexpr int kPileCell = 8;      // counters per (position, pool) cell of the window table: base 0..3 + 4 * strand
constexpr int kPileSpan = 512;    // longest match block whose reference stretch is staged in shared memory (longer reads: global loads)
constexpr int kPileRefWords = (kPileTile + 2 * kPileSpan + 16) / 4;

// the 4-bit code whose character (bam_nt16_rev_table) is the upper-cased reference character, 0xFF when there is none:
// read and reference characters are equal exactly when the codes are
__device__ __forceinline__ int ref_nt16(int up)
{
	switch (up) {
	case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; case 'N': return 15; case '=': return 0;
	case 'M': return 3; case 'R': return 5; case 'S': return 6; case 'V': return 7; case 'W': return 9; case 'Y': return 10;
	case 'H': return 11; case 'K': return 12; case 'D': return 13; case 'B': return 14;
	default: return 0xFF;
	}
}

// 0x01 in every byte b of a word with lo <= b < hi (any integers)
__device__ __forceinline__ uint32_t byte_range01(int lo, int hi)
{
	lo = max(lo, 0); hi = min(hi, 4);
	if (lo >= hi) return 0u;
	const uint32_t upto_hi = hi >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - hi)));
	return upto_hi & (0x01010101u << (8 * lo));
}

// hist index of counter e at tile offset x.  Counter-major; inside every block of 128 positions the column is transposed
// ([x & 3][x >> 2]) so that the 32 lanes of a warp — which hold four consecutive bases each — hit 32 different banks.
__device__ __forceinline__ int pile_slot(int e, int x) { return e * kPileTile + (x & ~127) + ((x & 3) << 5) + ((x >> 2) & 31); }
// BTI of a 4-bit base code: the table "0 0 1 0 2 0 1 0 3 0 1 0 2 0 1 0" is the index of the lowest set bit (0 for code 0)
__device__ __forceinline__ int nt16_bti_ffs(int c) { return max(__ffs(c) - 1, 0); }
// eight 4-bit bases of a 32-bit word of BAM sequence in linear order (base k in bits 4k .. 4k+3): BAM keeps the first base
// of a byte in the high nibble
__device__ __forceinline__ uint32_t linear_nibbles(uint32_t v) { return ((v & 0x0F0F0F0Fu) << 4) | ((v >> 4) & 0x0F0F0F0Fu); }

// first alignment of [lo, hi) whose position is >= key, by the whole warp: 32 probes per round instead of one
__device__ __forceinline__ uint32_t lower_bound_warp(const crispgpu_alnrec* rec, uint32_t lo, uint32_t hi, int key, int lane)
{
	while (hi - lo > 32u) {
		const uint32_t step = (hi - lo + 31u) >> 5;
		const uint32_t at = lo + (uint32_t)lane * step;
		const bool below = at < hi && rec[at].pos < key;       // monotone in the lane: a prefix of the lanes says yes
		const int nb = __popc(__ballot_sync(0xffffffffu, below));
		if (nb == 0) return lo;
		const uint32_t nlo = lo + (uint32_t)(nb - 1) * step + 1u, nhi = lo + (uint32_t)nb * step;   // rec[nlo-1] < key <= rec[nhi] (if any)
		lo = nlo;
		hi = nhi < hi ? nhi : hi;
	}
	const bool below = lo + lane < hi && rec[lo + lane].pos < key;
	return lo + (uint32_t)__popc(__ballot_sync(0xffffffffu, below));
}

// One alignment of phase B: the lane's four bases (one code per byte in `sq`, qualities in `q`) against the tile.
template <bool STAGED, class RefQuad>
__device__ __forceinline__ void pile_count_quad(uint32_t q, uint32_t h, int p0, int i0, int lo_i, int hi_i, bool refs, int st, int t0, uint32_t minq4,
                                                uint32_t* hist, RefQuad&& ref_quad)
{
	// memory byte 0 = (base 0 high nibble, base 1 low nibble), byte 1 = (base 2, base 3): one code per byte
	const uint32_t sq = __byte_perm((h & 0xF0F0u) >> 4, h & 0x0F0Fu, 0x5140);
	const uint32_t rq = ref_quad(p0);
	const uint32_t x = sq ^ rq;
	const uint32_t y = (x | (x >> 4)) & 0x0F0F0F0Fu;
	const uint32_t neq = ((y + 0x0F0F0F0Fu) >> 4) & 0x01010101u;                        // differs from the reference
	const uint32_t qok = ((((q | 0x80808080u) - minq4) | q) >> 7) & 0x01010101u;       // quality >= MINQ (MINQ <= 127)
	const uint32_t valid = (i0 >= lo_i && i0 + 4 <= hi_i) ? 0x01010101u : byte_range01(lo_i - i0, hi_i - i0);
	// exceptions: with `refs` everything but (equal, quality ok); without, only the countable non-reference bases
	uint32_t exc = valid & (refs ? (neq | (qok ^ 0x01010101u)) : (neq & qok));
	while (exc) {
		const int bit = __ffs(exc) - 1;
		exc &= exc - 1;
		const int x0 = p0 + (bit >> 3) - t0;
		const int rc = (rq >> bit) & 255;
		if (refs && rc != 0xFF) atomicSub(&hist[pile_slot(nt16_bti_ffs(rc) + st, x0)], 1u);
		if ((neq >> bit) & (qok >> bit) & 1u) atomicAdd(&hist[pile_slot(nt16_bti_ffs((sq >> bit) & 15) + st, x0)], 1u);
	}
}

// One CTA per (tile of kPileTile positions, pool): the alignments overlapping the tile are walked once, 32 at a time per warp.
//
// Phase A, one LANE per alignment: the read's mismatches against the reference (the read filter needs them before any base
// is counted), eight bases per step — the read's 4-bit codes XOR the reference stretch of the tile, which is staged in shared
// memory as a stream of 4-bit codes; "differs" and "is N" are nibble masks, the count is a popc.  The lane leaves the record
// and the filter outcome in the warp's slot of shared memory.
// Phase B: an alignment that passed is counted by its own lane as well, eight bases per step (the 4-bit codes again, the eight
// quality bytes as two words, "differs", "quality >= MINQ" and "inside the tile" as eight-bit masks) — or, when the tile's
// reference stretch holds a character outside BAM's alphabet or a read is longer than the staged stretch, by the whole warp,
// each lane holding four consecutive bases as the bytes of a word.  A read that may count reference
// bases adds +1 / -1 to a per-strand coverage difference array at its first / past-last position inside the tile; only its
// EXCEPTIONS touch the counters: a base that differs from the reference (counted under its own allele when its quality
// passes) or a reference base below MINQ takes one off the reference allele's counter.  After the walk a prefix sum of the
// difference arrays gives, per position and strand, the reads that cover it, which are added to the reference allele's
// counter — so the common case (base equals the reference, quality passes) costs no atomic at all.
// The table of 8 counters per position (base + 4 strand) leaves as full 32-byte cells of the window table
// win[position][pool][8].  An alignment is loaded by at most two tiles.
template <bool STAGED>
__device__ __forceinline__ void pile_tile_body(const PileReads& rd, const PileWindow& w, uint32_t* __restrict__ win, uint32_t* hist, int (*cov)[kPileTile + 8],
                                               uint32_t* sref, uint32_t* sref4, uint4* srec, int* stot, int max_span, bool fast)
{
	const int pool = blockIdx.y, t0 = w.w0 + blockIdx.x * kPileTile, t1 = min(t0 + kPileTile, w.w0 + w.W);
	const int s0 = t0 - kPileSpan - 4;   // reference position of the first staged code
	// reference code at position p / codes of p0 .. p0+3, one per byte
	auto ref_code1 = [&](int p) -> int { return STAGED ? (int)((sref[(p - s0) >> 2] >> (8 * ((p - s0) & 3))) & 255u) : ref_nt16(ref_at(w, p)); };
	auto ref_quad = [&](int p0) -> uint32_t {
		if (STAGED) {
			const int j = p0 - s0;
			return __funnelshift_r(sref[j >> 2], sref[(j >> 2) + 1], 8 * (j & 3));
		}
		uint32_t v = 0;
#pragma unroll
		for (int b = 0; b < 4; b++) v |= (uint32_t)ref_nt16(ref_at(w, p0 + b)) << (8 * b);
		return v;
	};
	const uint32_t minq4 = (uint32_t)(w.min_q < 0 ? 0 : (w.min_q > 127 ? 127 : w.min_q)) * 0x01010101u;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	const uint32_t r0 = rd.pool_off[pool], r1 = rd.pool_off[pool + 1];
	const uint32_t lo = lower_bound_warp(rd.rec, r0, r1, t0 - max_span + 1, lane), hi = lower_bound_warp(rd.rec, r0, r1, t1, lane);
	uint4* myrec = srec + 32 * warp;
	for (uint32_t g = lo + 32u * warp; g < hi; g += 32u * nw) {
		// ---- phase A: lane = alignment g + lane ----
		const uint32_t mine = g + lane;
		uint4 packed = make_uint4(0u, 0u, 0u, 0u);
		if (mine < hi) {
			const crispgpu_alnrec a = rd.rec[mine];
			const int ml = a.mlen, clip = a.clip;
			int mm = 0;
			if (STAGED && fast) {
				const uint32_t* sw = reinterpret_cast<const uint32_t*>(rd.seq4 + (a.base_off >> 1)) + (clip >> 3);
				const int sh = 4 * (clip & 7), lastw = ((clip + ml - 1) >> 3) - (clip >> 3);
				// four steps per trip: their five sequence words are loaded together (five loads in flight per lane)
				for (int tb = 0; 8 * tb < ml; tb += 4) {
					uint32_t wv[5];
#pragma unroll
					for (int j = 0; j < 5; j++) wv[j] = (tb + j <= lastw) ? __ldg(sw + tb + j) : 0u;
#pragma unroll
					for (int j = 0; j < 4; j++) {
						const int t = tb + j;
						if (8 * t >= ml) break;
						const uint32_t s8 = __funnelshift_r(linear_nibbles(wv[j]), linear_nibbles(wv[j + 1]), sh);
						const int jn = a.pos + 8 * t - s0;
						const uint32_t r8 = __funnelshift_r(sref4[jn >> 3], sref4[(jn >> 3) + 1], 4 * (jn & 7));
						uint32_t x = s8 ^ r8;
						x |= x >> 1; x |= x >> 2;
						const uint32_t isn = s8 & (s8 >> 1) & (s8 >> 2) & (s8 >> 3);
						uint32_t m = x & ~isn & 0x11111111u;
						const int left = ml - 8 * t;
						if (left < 8) m &= (1u << (4 * left)) - 1u;
						mm += __popc(m);
					}
				}
			} else {
				for (int k = 0; k < ml; k++) {
					const uint32_t i = a.base_off + clip + k;
					const int code = (rd.seq4[i >> 1] >> ((i & 1) ? 0 : 4)) & 15;
					mm += (code != 15 && code != ref_code1(a.pos + k)) ? 1 : 0;
				}
			}
			const int maxmm = ml / 40 + 2;   // MAX_MM = alignedbases/40 + 2 (no gaps in this version)
			const bool mq = a.mapq >= w.min_m;
			int fl = (mq && mm <= maxmm ? 1 : 0) | (mq && mm <= maxmm - 1 ? 2 : 0);
			rd.flags[mine] = (uint8_t)fl;   // tiles that share the read write the same value
			if (a.pos + ml <= t0) fl = 0;   // nothing of it inside this tile
			if (STAGED && fast) {
				// ---- phase B, fast form: the lane counts its own alignment, eight bases per step ----
				if (fl & 1) {
					const bool refs = (fl & 2) != 0;   // reference bases of this read count
					const int st = a.strand ? 4 : 0;
					// indices of the match block that lie inside the tile
					const int k_lo = max(t0 - a.pos, 0), k_hi = min(ml, t1 - a.pos);
					if (refs) {
						int* cs = cov[st >> 2];
						atomicAdd(cs + (a.pos + k_lo - t0), 1);
						atomicAdd(cs + (a.pos + k_hi - t0), -1);
					}
					const uint32_t* sw = reinterpret_cast<const uint32_t*>(rd.seq4 + (a.base_off >> 1)) + (clip >> 3);
					const int sh = 4 * (clip & 7), lastw = ((clip + ml - 1) >> 3) - (clip >> 3);
					const uint32_t* qw = reinterpret_cast<const uint32_t*>(rd.qual + a.base_off) + (clip >> 2);
					const int qsh = 8 * (clip & 3), lastq = ((clip + ml - 1) >> 2) - (clip >> 2);
					const int t_first = k_lo >> 3, t_end = (k_hi + 7) >> 3;   // steps that touch the tile
					for (int tb = t_first; tb < t_end; tb += 4) {
						// four steps per trip: five sequence words and nine quality words loaded together
						uint32_t wv[5], qv[9];
#pragma unroll
						for (int j = 0; j < 5; j++) wv[j] = (tb + j <= lastw) ? __ldg(sw + tb + j) : 0u;
#pragma unroll
						for (int j = 0; j < 9; j++) qv[j] = (2 * tb + j <= lastq) ? __ldg(qw + 2 * tb + j) : 0u;
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const int t = tb + j;
							if (t >= t_end) break;
							const uint32_t s8 = __funnelshift_r(linear_nibbles(wv[j]), linear_nibbles(wv[j + 1]), sh);
							const uint32_t qa = __funnelshift_r(qv[2 * j], qv[2 * j + 1], qsh), qb = __funnelshift_r(qv[2 * j + 1], qv[2 * j + 2], qsh);
							const int jn = a.pos + 8 * t - s0;
							const uint32_t r8 = __funnelshift_r(sref4[jn >> 3], sref4[(jn >> 3) + 1], 4 * (jn & 7));
							uint32_t x = s8 ^ r8;
							x |= x >> 1; x |= x >> 2;
							x &= 0x11111111u;                                        // nibble k: base k differs from the reference
							x = (x | (x >> 3)) & 0x03030303u; x = (x | (x >> 6)) & 0x000F000Fu; x = (x | (x >> 12)) & 0xFFu;   // one bit per base
							const uint32_t ma = ((((qa | 0x80808080u) - minq4) | qa) >> 7) & 0x01010101u, mb = ((((qb | 0x80808080u) - minq4) | qb) >> 7) & 0x01010101u;
							const uint32_t qok = ((ma * 0x01020408u) >> 24) | (((mb * 0x01020408u) >> 24) << 4);   // bit k: quality of base k >= MINQ
							const int lo = max(k_lo - 8 * t, 0), hi8 = min(k_hi - 8 * t, 8);
							const uint32_t valid = ((1u << hi8) - 1u) & ~((1u << lo) - 1u);
							uint32_t exc = valid & (refs ? (x | (qok ^ 0xFFu)) : (x & qok));
							while (exc) {
								const int b = __ffs(exc) - 1;
								exc &= exc - 1;
								const int x0 = a.pos + 8 * t + b - t0;
								if (refs) atomicSub(&hist[pile_slot(nt16_bti_ffs((r8 >> (4 * b)) & 15) + st, x0)], 1u);
								if ((x >> b) & (qok >> b) & 1u) atomicAdd(&hist[pile_slot(nt16_bti_ffs((s8 >> (4 * b)) & 15) + st, x0)], 1u);
							}
						}
					}
				}
				fl = 0;   // done: nothing left for the warp-per-alignment form below
			}
			packed = make_uint4((uint32_t)a.pos, a.base_off, (uint32_t)a.clip | ((uint32_t)a.mlen << 16), (uint32_t)fl | (a.strand ? 4u : 0u));
		}
		__syncwarp();
		myrec[lane] = packed;
		__syncwarp();
		// ---- phase B: the warp takes the alignments that passed one after the other ----
		unsigned todo = __ballot_sync(0xffffffffu, packed.w & 1u);
		uint32_t q_n = 0, h_n = 0;
		auto issue = [&](int j) {   // the first 128 bases of alignment j: loads in flight while the previous one is counted
			const uint4 rj = myrec[j];
			const int i0 = 4 * lane;
			if (i0 < (int)((rj.z & 0xffffu) + (rj.z >> 16))) {
				q_n = __ldg(reinterpret_cast<const uint32_t*>(rd.qual + rj.y + i0));
				h_n = __ldg(reinterpret_cast<const uint16_t*>(rd.seq4 + (rj.y >> 1) + (i0 >> 1)));
			}
		};
		if (todo) issue(__ffs(todo) - 1);
		while (todo) {
			const int j = __ffs(todo) - 1;
			todo &= todo - 1;
			const uint4 rj = myrec[j];
			const uint32_t q0 = q_n, h0 = h_n;
			if (todo) issue(__ffs(todo) - 1);
			const int pos = (int)rj.x, clip = (int)(rj.z & 0xffffu), ml = (int)(rj.z >> 16), end = clip + ml, st = (int)(rj.w & 4u);
			const bool refs = (rj.w & 2u) != 0;   // reference bases of this read count
			// base indices of the read that lie inside the match block and inside the tile
			const int lo_i = clip + max(t0 - pos, 0), hi_i = min(end, clip + t1 - pos);
			if (refs && lane == 0) {
				int* cs = cov[st >> 2];
				atomicAdd(cs + (pos + lo_i - clip - t0), 1);
				atomicAdd(cs + (pos + hi_i - clip - t0), -1);
			}
			{
				const int i0 = 4 * lane;
				if (i0 < hi_i && i0 + 4 > lo_i) pile_count_quad<STAGED>(q0, h0, pos + i0 - clip, i0, lo_i, hi_i, refs, st, t0, minq4, hist, ref_quad);
			}
			for (int c = 128; c < hi_i; c += 128) {   // reads longer than one pass of the warp
				const int i0 = c + 4 * lane;
				if (i0 >= hi_i || i0 + 4 <= lo_i) continue;
				const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(rd.qual + rj.y + i0));
				const uint32_t h = __ldg(reinterpret_cast<const uint16_t*>(rd.seq4 + (rj.y >> 1) + (i0 >> 1)));
				pile_count_quad<STAGED>(q, h, pos + i0 - clip, i0, lo_i, hi_i, refs, st, t0, minq4, hist, ref_quad);
			}
		}
	}
	__syncthreads();
	// coverage: prefix sums of the two difference arrays — warp k scans a quarter of one strand, the quarters' totals are added
	// when the sums go to the reference allele's counters
	{
		const int s = warp & 1, qd = warp >> 1, x0 = qd * (kPileTile / 4);
		int run = 0;
		for (int x = x0 + lane; x < x0 + kPileTile / 4; x += 32) {
			int v = cov[s][x];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
			v += run;
			run = __shfl_sync(0xffffffffu, v, 31);
			cov[s][x] = v;
		}
		if (lane == 0) stot[s * 4 + qd] = run;
	}
	__syncthreads();
	for (int k = threadIdx.x; k < 2 * (t1 - t0); k += blockDim.x) {
		const int s = k & 1, x = k >> 1, qd = x / (kPileTile / 4);
		int v = cov[s][x];
		for (int d = 0; d < qd; d++) v += stot[s * 4 + d];
		const int rc = ref_code1(t0 + x);
		if (rc != 0xFF && v) hist[pile_slot(nt16_bti_ffs(rc) + 4 * s, x)] += (uint32_t)v;
	}
	__syncthreads();
	// 32-byte cells: two 16-byte stores per cell, consecutive positions 32 * P bytes apart
	for (int k = threadIdx.x; k < (t1 - t0) * 2; k += blockDim.x) {
		const int x = k >> 1, h = k & 1;
		const uint4 v = make_uint4(hist[pile_slot(4 * h, x)], hist[pile_slot(4 * h + 1, x)], hist[pile_slot(4 * h + 2, x)], hist[pile_slot(4 * h + 3, x)]);
		*reinterpret_cast<uint4*>(win + ((size_t)(t0 - w.w0 + x) * rd.P + pool) * kPileCell + 4 * h) = v;
	}
}

__global__ void __launch_bounds__(256) k_pile_tile(PileReads rd, PileWindow w, uint32_t* __restrict__ win)
{
	__shared__ uint32_t hist[kPileTile * kPileCell];
	__shared__ int cov[2][kPileTile + 8];
	__shared__ uint32_t sref[kPileRefWords];        // reference codes, one per byte (0xFF: a character no read base can equal)
	__shared__ uint32_t sref4[kPileRefWords / 2 + 2];  // the same as a stream of 4-bit codes
	__shared__ uint4 srec[256];                     // per warp: the 32 alignments of its group (position, base offset, clip | length, flags)
	__shared__ int stot[8];
	if (rd.state[0]) return;
	const int max_span = rd.state[1];
	const bool staged = max_span <= kPileSpan;
	const int t0 = w.w0 + blockIdx.x * kPileTile;
	const int s0 = t0 - kPileSpan - 4;
	for (int k = threadIdx.x; k < kPileTile * kPileCell; k += blockDim.x) hist[k] = 0;
	for (int k = threadIdx.x; k < 2 * (kPileTile + 8); k += blockDim.x) (&cov[0][0])[k] = 0;
	int odd = 0;
	if (staged)
		for (int k = threadIdx.x; k < kPileRefWords; k += blockDim.x) {
			uint32_t v = 0;
#pragma unroll
			for (int b = 0; b < 4; b++) {
				const int code = ref_nt16(ref_at(w, s0 + 4 * k + b));
				odd |= code == 0xFF;
				v |= (uint32_t)code << (8 * b);
			}
			sref[k] = v;
		}
	const bool fast = staged && !__syncthreads_or(odd);   // (also the barrier after the zero fill and the staging)
	if (!staged) __syncthreads();
	if (fast) {
		for (int k = threadIdx.x; k < kPileRefWords / 2; k += blockDim.x) {
			const uint32_t a = sref[2 * k], b = sref[2 * k + 1];
			sref4[k] = (a & 0xFu) | ((a >> 4) & 0xF0u) | ((a >> 8) & 0xF00u) | ((a >> 12) & 0xF000u) |
			           ((b & 0xFu) << 16) | ((b << 12) & 0xF00000u) | ((b << 8) & 0xF000000u) | ((b << 4) & 0xF0000000u);
		}
		if (threadIdx.x < 2) sref4[kPileRefWords / 2 + threadIdx.x] = 0;
		__syncthreads();
	}
	if (staged) pile_tile_body<true>(rd, w, win, hist, cov, sref, sref4, srec, stot, max_span, fast);
	else pile_tile_body<false>(rd, w, win, hist, cov, sref, sref4, srec, stot, max_span, fast);
}

// sort_allelecounts (allelecounts.c:134-183, IUPAC off) on the eight per-allele read totals: reference allele swapped to
// the front, then the exchange sort `if (ct[i] < ct[j]) swap` over i = 1..6, j = i+1..7 — the same comparisons in the
// same order, so ties fall the same way
__device__ __forceinline__ void sort_alleles(int ref, int (&ct)[8], int (&al)[8])
{
	if (ref >= 1 && ref < 8) {
#pragma unroll
		for (int a = 1; a < 8; a++)
			if (a == ref) { const int tc = ct[0]; ct[0] = ct[a]; ct[a] = tc; const int ta = al[0]; al[0] = al[a]; al[a] = ta; }
	}
#pragma unroll
	for (int i = 1; i < 7; i++)
#pragma unroll
		for (int j = i + 1; j < 8; j++)
			if (ct[i] < ct[j]) { const int tc = ct[i]; ct[i] = ct[j]; ct[j] = tc; const int ta = al[i]; al[i] = al[j]; al[j] = ta; }
}

// Candidate test of every window position (variantcalls.c:47,76-92): one CTA per 32 consecutive positions, four per warp,
// lanes over pools; the potential-variant score adds 2 per (non-reference allele, pool) with >= 3 reads, and the read count
// for diploid pools with >= 1.  A position is a candidate when the score reaches 2 on an A/C/G/T reference base: one bit per
// position, bits[x / 32].
__global__ void __launch_bounds__(256) k_pile_screen(const uint32_t* __restrict__ win, PileWindow w, int P, const int32_t* __restrict__ ploidy, uint32_t* __restrict__ bits)
{
	__shared__ uint32_t word;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) word = 0;
	__syncthreads();
	uint32_t mask = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int x = blockIdx.x * 32 + warp * 4 + k;
		if (x >= w.W) break;   // warp-uniform
		const int ref = ref_code(ref_at(w, w.w0 + x));
		int pot = 0;
		for (int i = lane; i < P; i += 32) {
			const uint4* c = reinterpret_cast<const uint4*>(win + ((size_t)x * P + i) * kPileCell);
			const uint4 f = __ldg(c), r = __ldg(c + 1);
			const int t[4] = {(int)(f.x + r.x), (int)(f.y + r.y), (int)(f.z + r.z), (int)(f.w + r.w)};
			const bool diploid = ploidy[i] == 2;
#pragma unroll
			for (int a = 0; a < 4; a++) {
				if (a == ref) continue;
				if (t[a] >= 3) pot += 2;
				else if (t[a] >= 1 && diploid) pot += t[a];
			}
		}
		pot = __reduce_add_sync(0xffffffffu, pot);
		if (pot >= 2 && ref < 4) mask |= 1u << (warp * 4 + k);
	}
	if (lane == 0 && mask) atomicOr(&word, mask);
	__syncthreads();
	if (threadIdx.x == 0) bits[blockIdx.x] = word;
}

// candidate positions in coordinate order: CTA b owns 64 words of the bit map (2048 positions); it counts the candidates
// before them, then every thread writes the candidates of its word
constexpr int kPileSelectWords = 64;
__global__ void __launch_bounds__(kPileSelectWords) k_pile_select(const uint32_t* __restrict__ bits, int n_words, int32_t* __restrict__ cand, int32_t* n_cand)
{
	__shared__ int part[2];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int w0 = blockIdx.x * kPileSelectWords;
	int before = 0;
	for (int k = threadIdx.x; k < w0; k += blockDim.x) before += __popc(__ldg(bits + k));
	before = __reduce_add_sync(0xffffffffu, before);
	const int wi = w0 + threadIdx.x;
	uint32_t mine = wi < n_words ? bits[wi] : 0u;
	int inc = __popc(mine);
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
	if (lane == 31) part[warp] = inc;   // the warp's candidates
	const int before_all = before;
	__shared__ int bef[2];
	if (lane == 0) bef[warp] = before_all;
	__syncthreads();
	int o = bef[0] + bef[1] + (warp ? part[0] : 0) + inc - __popc(mine);
	while (mine) { const int bit = __ffs(mine) - 1; mine &= mine - 1; cand[o++] = wi * 32 + bit; }
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_cand = bef[0] + bef[1] + part[0] + part[1];
}

// exclusive prefix sum of n 32-bit counts into off[0..n] by one CTA (the arrays scanned here hold a few 10^5 cells)
__global__ void k_pile_scan(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ off, size_t n)
{
	__shared__ unsigned long long sums[1024];
	const int T = blockDim.x, tid = threadIdx.x;
	const size_t chunk = (n + T - 1) / T, a = min(n, (size_t)tid * chunk), b = min(n, a + chunk);
	unsigned long long c = 0;
	for (size_t x = a; x < b; x++) c += cnt[x];
	sums[tid] = c;
	__syncthreads();
	for (int d = 1; d < T; d <<= 1) {
		const unsigned long long v = tid >= d ? sums[tid - d] : 0;
		__syncthreads();
		sums[tid] += v;
		__syncthreads();
	}
	unsigned long long o = sums[tid] - c;
	for (size_t x = a; x < b; x++) { off[x] = (uint32_t)o; o += cnt[x]; }
	if (tid == T - 1) off[n] = (uint32_t)sums[tid];
}

struct PileSites {   // the candidate sites of the window as the engine's batch arrays (device memory)
	int n;
	const int32_t* cand;     // [n] window offset of each candidate
	int32_t* tid; int32_t* pos; uint8_t* refbase; uint8_t* maxvec_al; int32_t* maxvec_ct; int32_t* counts; int32_t* indcounts;
	uint32_t* cell_reads;    // [n P] counted reads per (site, pool) -> read_off by k_pile_scan
	uint32_t* read_off; uint16_t* reads;
	uint32_t* alt_len;       // [n 5 P] alt-read list lengths per (site, slot, pool) -> alt_cell_off
	uint32_t* alt_cell_off; uint32_t* alt_off; crispgpu_altread* alt_reads;
	int32_t* readdepths;     // [n P]
	int32_t* mqcounts;       // [n 4]
	int32_t* filtered;       // [n 2]
	double* stats;           // [n P 16] or null: sum of 10^(-q/10) over the counted reads per allele + 8 strand (variant->stats; --EM 0)
	double* tstats;          // [n 16]
};

// One CTA per candidate: its cells of the window table become the dense per-pool rows of the batch (indel and
// bidirectional entries zero), the site totals, the sorted allele vector and the per-cell read counts.
__global__ void __launch_bounds__(128) k_pile_gather_counts(PileSites ps, PileWindow w, int P, const uint32_t* __restrict__ win)
{
	__shared__ int tot[kPileCell];
	for (int c = blockIdx.x; c < ps.n; c += gridDim.x) {
		const int x = ps.cand[c];
		if (threadIdx.x < kPileCell) tot[threadIdx.x] = 0;
		__syncthreads();
		int mine[kPileCell] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int j = threadIdx.x; j < P; j += blockDim.x) {
			const uint4* cell = reinterpret_cast<const uint4*>(win + ((size_t)x * P + j) * kPileCell);
			const uint4 f = cell[0], r = cell[1];
			const int v[kPileCell] = {(int)f.x, (int)f.y, (int)f.z, (int)f.w, (int)r.x, (int)r.y, (int)r.z, (int)r.w};
			int4* row = reinterpret_cast<int4*>(ps.indcounts + ((size_t)c * P + j) * kNC);
			row[0] = make_int4(v[0], v[1], v[2], v[3]);
			row[1] = make_int4(0, 0, 0, 0);
			row[2] = make_int4(v[4], v[5], v[6], v[7]);
			row[3] = make_int4(0, 0, 0, 0);
			row[4] = make_int4(0, 0, 0, 0);
			row[5] = make_int4(0, 0, 0, 0);
			int t = 0;
#pragma unroll
			for (int e = 0; e < kPileCell; e++) { mine[e] += v[e]; t += v[e]; }
			ps.cell_reads[(size_t)c * P + j] = (uint32_t)t;
		}
#pragma unroll
		for (int e = 0; e < kPileCell; e++) {
			const int s = __reduce_add_sync(0xffffffffu, mine[e]);
			if ((threadIdx.x & 31) == 0 && s) atomicAdd(&tot[e], s);
		}
		__syncthreads();
		if (threadIdx.x < kNC) {
			const int e = threadIdx.x;
			ps.counts[c * kNC + e] = (e < 4) ? tot[e] : (e >= 8 && e < 12) ? tot[e - 4] : 0;
		}
		if (threadIdx.x == 0) {
			const int ref = ref_code(ref_at(w, w.w0 + x));
			int ct[8], al[8];
#pragma unroll
			for (int a = 0; a < 8; a++) { ct[a] = a < 4 ? tot[a] + tot[a + 4] : 0; al[a] = a; }
			sort_alleles(ref, ct, al);
#pragma unroll
			for (int a = 0; a < 8; a++) { ps.maxvec_al[c * 8 + a] = (uint8_t)al[a]; ps.maxvec_ct[c * 8 + a] = ct[a]; }
			ps.tid[c] = w.tid; ps.pos[c] = w.w0 + x; ps.refbase[c] = (uint8_t)ref;
		}
		if (threadIdx.x < 4) ps.mqcounts[c * 4 + threadIdx.x] = 0;
		if (threadIdx.x < 2) ps.filtered[c * 2 + threadIdx.x] = 0;
		__syncthreads();
	}
}

// One warp per (candidate, pool): the reads of the pool that cover the position, in BAM order.  PASS 0 measures the
// alt-read lists; PASS 1 writes the likelihood stream, the alt-read lists, the read depth, the mapping-quality classes
// and the two filtered-read counters.
template <int PASS>
__global__ void k_pile_gather_reads(PileReads rd, PileWindow w, PileSites ps)
{
	const int lane = threadIdx.x & 31, P = rd.P;
	const int max_span = rd.state[1];
	const size_t cells = (size_t)ps.n * P;
	for (size_t cell = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); cell < cells; cell += (size_t)gridDim.x * (blockDim.x >> 5)) {
		const int c = (int)(cell / P), j = (int)(cell - (size_t)c * P);
		const int p = ps.pos[c];
		const int refc = ref_at(w, p);
		const uint32_t r0 = rd.pool_off[j], r1 = rd.pool_off[j + 1];
		const uint32_t lo = lower_bound_warp(rd.rec, r0, r1, p - max_span + 1, lane), hi = lower_bound_warp(rd.rec, r0, r1, p + 1, lane);
		int slot_al[5];
#pragma unroll
		for (int k = 0; k < 5; k++) slot_al[k] = ps.maxvec_ct[c * 8 + k + 1] >= 3 ? ps.maxvec_al[c * 8 + k + 1] : -1;
		uint32_t nread = 0, nalt[5] = {0, 0, 0, 0, 0};
		int depth = 0, mq[4] = {0, 0, 0, 0}, filt[2] = {0, 0};
		double ep_sum = 0.0;   // lane e < 16: error-probability sum of the reads counted under allele (e & 7), strand (e >> 3), in read order
		const uint32_t rbase = PASS ? ps.read_off[cell] : 0;
		for (uint32_t rb = lo; rb < hi; rb += 32) {
			const uint32_t r = rb + lane;
			int b = -1, q = 0;
			bool covers = false, letter = false;   // letter: the base is the allele's own character (an 'N' counts as allele 0 but is
			                                       // no alt read of it: calculate_uniquereads compares characters, allelecounts.c:55)
			crispgpu_alnrec a{};
			if (r < hi) {
				a = rd.rec[r];
				const int k = p - a.pos;
				covers = k >= 0 && k < (int)a.mlen;
				if (covers) {
					const uint32_t i = a.base_off + a.clip + k;
					q = rd.qual[i];
					const int code = (rd.seq4[i >> 1] >> ((i & 1) ? 0 : 4)) & 15;
					const int fl = rd.flags[r];
					const bool mqok = a.mapq >= w.min_m;
					if (PASS) {
						// variant.c:288-289: MAX_MM test first, then the base quality
						if (mqok && !(fl & 1)) filt[0]++;
						else if (q < w.min_q && mqok) filt[1]++;
					}
					if (q >= w.min_q && (fl & 1) && !(nt16_char(code) == refc && !(fl & 2))) { b = nt16_bti(code); letter = code == (1 << b); }
				}
			}
			if (PASS) {
				const unsigned cm = __ballot_sync(0xffffffffu, covers);
				depth += __popc(cm);
				if (covers) { const int m = a.mapq; mq[m < 10 ? 0 : (m >= 40 ? 3 : m / 20 + 1)]++; }
				const unsigned bm = __ballot_sync(0xffffffffu, b >= 0);
				if (b >= 0) ps.reads[rbase + nread + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)(((q + 33) & 0xff) | ((b & 3) << 8) | ((a.strand ? 1 : 0) << 10));
				nread += __popc(bm);
				if (ps.stats) {
					// ep = 0.1^(quality/10) added in read order (variant.c:318-319): lane l's read goes to the lane that owns its class
					const double ep = b >= 0 ? pow(0.1, (double)q / 10) : 0.0;
					const int cls = b >= 0 ? b + (a.strand ? 8 : 0) : -1;
					for (unsigned left = bm; left;) {
						const int l = __ffs(left) - 1;
						left &= left - 1;
						const double e = __shfl_sync(0xffffffffu, ep, l);
						if (__shfl_sync(0xffffffffu, cls, l) == lane) ep_sum += e;
					}
				}
			}
#pragma unroll
			for (int k = 0; k < 5; k++) {
				if (slot_al[k] < 0) continue;   // warp-uniform
				const unsigned am = __ballot_sync(0xffffffffu, letter && b == slot_al[k]);
				if (PASS && letter && b == slot_al[k]) {
					crispgpu_altread e;
					e.pool = (uint16_t)j;
					e.offset = (uint16_t)(a.clip + (p - a.pos));   // l1 + delta: index of the base in the read
					e.aligned = a.mlen;
					e.strand = a.strand ? 1 : 0;
					e.pad = 0;
					ps.alt_reads[ps.alt_cell_off[((size_t)c * 5 + k) * P + j] + nalt[k] + __popc(am & ((1u << lane) - 1u))] = e;
				}
				nalt[k] += __popc(am);
			}
		}
		if (!PASS) {
			if (lane < 5) {
				uint32_t v = 0;
#pragma unroll
				for (int k = 0; k < 5; k++) if (lane == k) v = nalt[k];
				ps.alt_len[((size_t)c * 5 + lane) * P + j] = v;
			}
		} else {
			if (lane == 0) ps.readdepths[cell] = depth;
			if (ps.stats && lane < 16) ps.stats[cell * 16 + lane] = ep_sum;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int t = __reduce_add_sync(0xffffffffu, mq[k]);
				if (lane == 0 && t) atomicAdd(&ps.mqcounts[c * 4 + k], t);
			}
#pragma unroll
			for (int k = 0; k < 2; k++) {
				const int t = __reduce_add_sync(0xffffffffu, filt[k]);
				if (lane == 0 && t) atomicAdd(&ps.filtered[c * 2 + k], t);
			}
		}
	}
}

// variant->tstats: the error-probability sums over all pools, added in pool order
__global__ void k_pile_tstats(PileSites ps, int P)
{
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ps.n * 16; k += gridDim.x * blockDim.x) {
		const int c = k >> 4, e = k & 15;
		double t = 0.0;
		for (int j = 0; j < P; j++) t += ps.stats[((size_t)c * P + j) * 16 + e];
		ps.tstats[k] = t;
	}
}

// alt_off[site * 5 + slot] = start of the slot's list (its pools follow one another)
__global__ void k_pile_alt_off(PileSites ps, int P)
{
	const int n5 = ps.n * 5;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= n5; k += gridDim.x * blockDim.x) ps.alt_off[k] = ps.alt_cell_off[(size_t)k * P];
}

}  // namespace crisp
