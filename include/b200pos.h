/* b200pos.h -- C ABI of libb200pos.so: B200-native attestation aggregation + LMD-GHOST fork choice.
 *
 * Drop-in boundary for the hot path of /root/reference/pos-evolution.md
 *     process_attestation (:722) -> bls.Aggregate / FastAggregateVerify -> get_head (:1102)
 * The reference is executable Python (pyspec) and has no FFI of its own; the functions below are
 * what a ctypes binding behind the pyspec names would call (INTEGRATION.md shows the stub).
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no C++/torch types; the caller owns every buffer; the library keeps only the
 *     registry / block-tree / latest-message tables it allocated on the device.
 *   - return value: 0 = ok, negative = B2_E* (environment / argument errors).  Cryptographic
 *     invalidity is DATA (ok_out[i] = 0, seg_status[i] != 0), never an error code.  On error no
 *     output buffer and no device table is modified (pos-evolution.md:1041).
 *   - functions without the _dev suffix take HOST pointers, are synchronous and do their own
 *     host<->device copies.  *_dev functions take DEVICE pointers plus a cudaStream_t (as void*)
 *     and only enqueue work (used by bench.py and by the multi-GPU path, where the u64 vote
 *     weights are all-reduced by NCCL between b2_vote_weights_dev and b2_head_from_votes_dev).
 *   - a b2_ctx is bound to one GPU and is not thread-safe (pyspec is single-threaded).
 *   - there is no CPU fallback: b2_init fails with B2_ENODEVICE when no sm_100 GPU is present.
 */
#ifndef B200POS_H
#define B200POS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_ctx b2_ctx;

enum {
    B2_OK = 0,
    B2_EINVAL = -1,    /* bad argument (null pointer, index out of range, missing table) */
    B2_ECUDA = -2,     /* CUDA runtime error; b2_last_error() has the text */
    B2_ENODEVICE = -3, /* no usable GPU */
    B2_ENOMEM = -4
};

/* status bits written by b2_g1_aggregate (why py_ecc's FastAggregateVerify would return False) */
enum { B2_PK_OK = 0, B2_PK_INVALID_KEY = 1, B2_PK_EMPTY = 2, B2_PK_INFINITY = 4, B2_PK_BAD_INDEX = 8 };

int b2_init(int device, b2_ctx** out);
void b2_destroy(b2_ctx* ctx);
const char* b2_last_error(b2_ctx* ctx);
int b2_sync(b2_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t b2_launch_count(b2_ctx* ctx);

/* ---- validator registry: BeaconState.validators[*].{pubkey, effective_balance} (pos-evolution.md:36-45, :354).
 * Decompresses and KeyValidates every pubkey ONCE (py_ecc does it on every FastAggregateVerify call) and
 * keeps 96-byte affine Montgomery records on the device.  flags: bit0 = active at the fork-choice epoch
 * (is_active_validator), bit1 = slashed.  pk_valid_out (may be NULL): 1 = KeyValidate passed. */
int b2_registry_load(b2_ctx* ctx, const uint8_t* pk48, const uint64_t* effective_balance, const uint8_t* flags,
                     uint64_t n_validators, uint8_t* pk_valid_out);
/* bls.KeyValidate (decodable, not infinity, in the r-torsion; eth2spec.utils.bls, used by process_deposit upstream) for n explicit
 * pubkeys; leaves the registry alone. */
int b2_key_validate(b2_ctx* ctx, const uint8_t* pk48, uint64_t n, uint8_t* valid_out);
/* refresh balances / flags only (process_effective_balance_updates, pos-evolution.md:122-133, changes them per epoch) */
int b2_registry_update_balances(b2_ctx* ctx, const uint64_t* effective_balance, const uint8_t* flags, uint64_t n_validators);

/* ---- committees and aggregation bits (get_beacon_committee :729, get_attesting_indices :745):
 * aggregate a has members[off[a] .. off[a+1]) (validator indices, committee order) and bit j of
 * bits[a*bits_stride ...] (little-endian within bytes, SSZ Bitlist order) selects member j. */

/* K2: per aggregate, sum of the selected pubkeys; out48 = compressed G1 (infinity encoding when empty). */
int b2_g1_aggregate(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride,
                    uint32_t n_agg, uint8_t* out48, uint8_t* status);

/* bls.Aggregate (eth2spec.utils.bls; named by BASELINE.json, called as get_aggregate_signature upstream), batched:
 * segment s = signatures [seg_off[s], seg_off[s+1]).  seg_status: 0 ok, 1 = a signature is undecodable
 * (py_ecc raises), 2 = empty segment (py_ecc raises).  No subgroup check, as in py_ecc. */
int b2_aggregate(b2_ctx* ctx, const uint8_t* sig96, const uint32_t* seg_off, uint32_t n_seg, uint8_t* out96, int32_t* seg_status);

/* bls.FastAggregateVerify inside is_valid_indexed_attestation (pos-evolution.md:736, :976), batched over
 * aggregates, pubkeys taken from the registry by index.  ok_out[a] in {0,1}; never an error for bad crypto. */
int b2_fast_aggregate_verify(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride,
                             const uint8_t* msg32, const uint8_t* sig96, uint32_t n_agg, uint8_t* ok_out);

/* pyspec-literal form: explicit compressed pubkeys, aggregate a uses pk48[pk_off[a] .. pk_off[a+1]).
 * Decompresses + KeyValidates every key on every call (what py_ecc does).  bls.Verify (:165) is the 1-key case. */
int b2_fast_aggregate_verify_pks(b2_ctx* ctx, const uint8_t* pk48, const uint32_t* pk_off, const uint8_t* msg32,
                                 const uint8_t* sig96, uint32_t n_agg, uint8_t* ok_out);

/* bls.SkToPk / bls.Sign, batched (sk = 8 little-endian u32 limbs, 0 < sk < r).  Signature i = sk[i] * H(msg[msg_idx[i]]). */
int b2_sk_to_pk(b2_ctx* ctx, const uint32_t* sk8, uint64_t n, uint8_t* pk48_out);
int b2_sign(b2_ctx* ctx, const uint32_t* sk8, const uint32_t* msg_idx, uint64_t n, const uint8_t* msg32, uint32_t n_msg,
            uint8_t* sig96_out);
/* H(m) compressed, for tests (hash_to_G2 with the POP ciphersuite tag) */
int b2_hash_to_g2(b2_ctx* ctx, const uint8_t* msg32, uint32_t n_msg, uint8_t* out96);

/* `hash` of the spec (SHA-256; pos-evolution.md:486, :522, :525), batched over n messages of msg_len bytes each */
int b2_sha256_batch(b2_ctx* ctx, const uint8_t* msgs, uint32_t msg_len, uint64_t n, uint8_t* out32);

/* SSZ wire form of n Attestations (pos-evolution.md:714-717, AttestationData :689-697): each is
 *   [u32 LE offset of aggregation_bits = 228][AttestationData 128 B][signature 96 B][Bitlist bytes, length given by the delimiter bit].
 * wire = the encodings back to back, woff u32[n+1] their byte offsets.  Outputs per attestation: the bit row (delimiter removed,
 * zero padded to bits_stride bytes -- the form b2_fast_aggregate_verify / b2_latest_messages_update take), its bit length, the
 * 128-byte AttestationData record (the input of b2_signing_roots) and the signature.
 * status_out: 0 ok; 1 malformed container; 2 empty Bitlist / no delimiter; 3 longer than max_bits (MAX_VALIDATORS_PER_COMMITTEE)
 * or than bits_stride allows.  A rejected attestation yields zero rows: malformed input is data, not an error. */
int b2_attestations_decode(b2_ctx* ctx, const uint8_t* wire, const uint32_t* woff, uint32_t n, uint32_t bits_stride, uint32_t max_bits,
                           uint8_t* bits_out, uint32_t* bit_len_out, uint8_t* data128_out, uint8_t* sig96_out, int32_t* status_out);

/* compute_signing_root(AttestationData, domain) (pattern of pos-evolution.md:163; containers :689-697, :219-221) for n attestations:
 * data128 = the 128-byte SSZ serialisation of each AttestationData; domain32 = one 32-byte domain for the batch, or one per
 * attestation when per_attestation_domain != 0.  SSZ merkleization (10 SHA-256 per attestation) runs on the GPU. */
int b2_signing_roots(b2_ctx* ctx, const uint8_t* data128, const uint8_t* domain32, int per_attestation_domain, uint32_t n, uint8_t* out32);

/* ---- committees: compute_committee / compute_shuffled_index (pos-evolution.md:495-534) for a whole epoch.
 * members_out[i] = active[compute_shuffled_index(i, n_active, seed)] (active == NULL: identity), so committee k of
 * `count` is members_out[n*k/count .. n*(k+1)/count).  SHA-256 and the swap-or-not rounds run on the GPU. */
int b2_shuffle_committees(b2_ctx* ctx, const uint8_t* seed32, const uint32_t* active, uint32_t n_active, uint32_t rounds,
                          uint32_t* members_out);
int b2_shuffle_committees_dev(b2_ctx* ctx, const uint8_t* d_seed32, const uint32_t* d_active, uint32_t n_active, uint32_t rounds,
                              uint32_t* d_members_out, void* stream);

/* ---- fork choice --------------------------------------------------------------------------------
 * Store.latest_messages (pos-evolution.md:901) lives on the device as (epoch, block index) per validator. */
int b2_latest_messages_reset(b2_ctx* ctx);
/* bulk load: has_msg[v] != 0 -> (epoch[v], block_idx[v]); equivocating[v] != 0 -> Store.equivocating_indices (:897) */
int b2_latest_messages_load(b2_ctx* ctx, const uint64_t* epoch, const uint32_t* block_idx, const uint8_t* has_msg,
                            const uint8_t* equivocating, uint64_t n_validators);
int b2_latest_messages_read(b2_ctx* ctx, uint64_t* epoch, uint32_t* block_idx, uint8_t* has_msg, uint64_t n_validators);
/* update_latest_messages (pos-evolution.md:1435-1441) for a batch of attestations in list order:
 * for each accepted aggregate a (accept[a] != 0, or accept == NULL) and each selected member i not equivocating:
 * set (target_epoch[a], block_idx[a]) iff i has no message or target_epoch[a] > stored epoch.  Order-exact. */
int b2_latest_messages_update(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride,
                              const uint64_t* target_epoch, const uint32_t* block_idx, const uint8_t* accept, uint32_t n_agg);

/* ---- participation flags and proposer-reward numerators: the bookkeeping loop of process_attestation (pos-evolution.md:738-754).
 * `which`: 0 = state.current_epoch_participation, 1 = previous.  For a batch of attestations in list order, each accepted
 * attestation a sets, for every selected member, the flags of flag_mask[a] (bit f = participation flag index f) that are not set
 * yet, and numerator_out[a] receives sum get_base_reward(index) * PARTICIPATION_FLAG_WEIGHTS[f] over the flags IT set
 * (get_base_reward = effective_balance / increment * base_reward_per_increment).  Order-exact for the whole batch. */
int b2_participation_load(b2_ctx* ctx, int which, const uint8_t* participation, uint64_t n_validators);
int b2_participation_read(b2_ctx* ctx, int which, uint8_t* participation_out, uint64_t n_validators);
int b2_participation_update(b2_ctx* ctx, int which, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride,
                            const uint8_t* flag_mask, const uint8_t* accept, uint32_t n_agg, uint64_t effective_balance_increment,
                            uint64_t base_reward_per_increment, uint64_t* numerator_out);

/* ---- FFG balance sums: the inputs of weigh_justification_and_finalization (pos-evolution.md:793-803, :817-837).
 * out4[0] = get_total_active_balance (sum of effective balances, registry flag bit0 = active in the current epoch),
 * out4[1] / out4[2] = get_total_balance(get_unslashed_participating_indices(state, flag_index, current / previous epoch)): validators
 *           that are active in that epoch (flag bit0 / bit2), not slashed (bit1) and carry `flag_index` in the participation table
 *           `which` = 0 / 1 (b2_participation_load; a table that was never loaded counts as all-zero),
 * out4[3] = total active AND unslashed balance (the essay's prose reading of get_total_active_balance, :809).
 * Raw sums: the caller applies max(EFFECTIVE_BALANCE_INCREMENT, .) as get_total_balance does. */
int b2_ffg_balances(b2_ctx* ctx, uint32_t flag_index, uint64_t* out4);

/* ---- fork-choice variants (pos-evolution.md:1411-1413, :1447-1461, :1549-1596):
 * b2_set_fork_choice_params: votes whose epoch is < min_vote_epoch do not count (vote expiry of the Goldfish / RLMD-GHOST family;
 *   0 = plain LMD-GHOST); exclude_slashed != 0 selects the v1.3 get_weight rule (slashed validators, registry flag bit1, do not count).
 * b2_on_attester_slashing: the validators present in BOTH sorted index lists become equivocating (Store.equivocating_indices)
 *   and stop counting; validity of the slashing itself (:1453-1457) is the caller's job. */
int b2_set_fork_choice_params(b2_ctx* ctx, uint64_t min_vote_epoch, int exclude_slashed);
/* How bls.FastAggregateVerify (is_valid_indexed_attestation, pos-evolution.md:736/:976) is evaluated for a BATCH of aggregates:
 *   mode 0 (default): e(PK_a, H(m_a)) * e(-g1, S_a) == 1 for every aggregate a (two Miller loops + one final exponentiation each);
 *   mode 1: random-linear-combination batches -- per group of 32 aggregates ONE equation
 *           prod_a e([r_a] PK_a, H(m_a)) * e(-g1, sum_a [r_a] S_a) == 1, r_a = 64-bit scalars derived by SHA-256 from `seed32`
 *           (the verifier's secret: pass fresh randomness), i.e. one Miller loop per aggregate plus one Miller loop and one final
 *           exponentiation per group; the members of a group whose equation fails are then verified one by one, so ok_out is
 *           the mode-0 vector (a wrong accept needs a 2^-63 coincidence).  Less work, longer critical path: for bulk verification.
 * Synchronises the device.  Applies to every FastAggregateVerify / epoch entry point of the context. */
int b2_set_verify_mode(b2_ctx* ctx, int mode, const uint8_t* seed32);
int b2_on_attester_slashing(b2_ctx* ctx, const uint32_t* indices_1, uint32_t n1, const uint32_t* indices_2, uint32_t n2);

/* Store.blocks (pos-evolution.md:898) as arrays in topological order (parent[b] < b, parent[0] ignored; block 0 =
 * store.justified_checkpoint.root).  leaf_viable[b]: the get_filtered_block_tree leaf test (:1104, prose :1121-1124). */
int b2_tree_load(b2_ctx* ctx, const uint32_t* parent, const uint64_t* slot, const uint8_t* root32, const uint8_t* leaf_viable,
                 uint32_t n_blocks);
/* get_latest_attesting_balance == get_weight (called at pos-evolution.md:1116) for EVERY block; boost_idx < 0: no boost */
int b2_get_weights(b2_ctx* ctx, int32_t boost_idx, uint64_t boost_score, uint64_t* weight_out);
/* get_head (pos-evolution.md:1102-1116): index of the head block */
int b2_get_head(b2_ctx* ctx, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score, uint32_t* head_idx_out);

/* ---- device-pointer entry points (asynchronous on `stream`) ------------------------------------------------------
 * The host cannot read what device pointers point to, so the checks the host entry points make before touching anything
 * (check_batch: monotone offsets, rows no longer than the bit row, member indices inside the registry, epochs < 2^32 - 1) are made by
 * the kernels themselves: an offending aggregate is skipped -- it fails verification (status B2_PK_BAD_INDEX), joins no LMD update,
 * its segment is reported undecodable -- nothing is read or written out of bounds, and the event is recorded in a guard word.
 * Precondition that cannot be checked on the device: off[n_agg] <= length of the members array.
 * b2_guard_flags: synchronises, returns and clears that word (bit0 member index >= n_validators, bit1 malformed offsets row,
 * bit2 target epoch does not fit 32 bits). */
enum { B2_GUARD_BAD_INDEX = 1, B2_GUARD_BAD_ROW = 2, B2_GUARD_BAD_EPOCH = 4 };
int b2_guard_flags(b2_ctx* ctx, uint32_t* flags_out);
int b2_aggregate_dev(b2_ctx* ctx, const uint8_t* d_sig96, const uint32_t* d_seg_off, uint32_t n_seg, uint64_t n_sig, uint8_t* d_out96,
                     int32_t* d_seg_status, void* stream);
int b2_fast_aggregate_verify_dev(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits,
                                 uint32_t bits_stride, const uint8_t* d_msg32, const uint8_t* d_sig96, uint32_t n_agg,
                                 uint8_t* d_ok_out, void* stream);
/* The gather stage of K2 alone, for measurement (north_star: ">= 60 % of HBM-read roofline on the pubkey gather"): for every aggregate
 * the XOR of the 24 record words of each selected member's pubkey record (get_attesting_indices + the pubkey list of
 * is_valid_indexed_attestation, pos-evolution.md:736/:745, without the additions).  form 1 = TMA-staged (cp.async.bulk -> shared
 * memory -> 128-bit LDS, the path b2_fast_aggregate_verify uses), form 0 = plain 128-bit global loads.  Both give the same words. */
int b2_gather_probe_dev(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride,
                        uint32_t n_agg, int form, uint32_t* d_checksum_out, void* stream);
/* One epoch of this rank in one call: bls.Aggregate per committee (segment a = signatures [off[a], off[a+1]), one per member, in
 * committee order) -> FastAggregateVerify of the aggregates -> update_latest_messages for the accepted ones.  The pubkey/hash half
 * of the verification is overlapped with the signature decompression on side streams. */
int b2_epoch_dev(b2_ctx* ctx, const uint8_t* d_sig96, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits,
                 uint32_t bits_stride, const uint8_t* d_msg32, const uint64_t* d_target_epoch, const uint32_t* d_block_idx, uint32_t n_agg,
                 uint64_t n_sig, uint8_t* d_agg_sig96, int32_t* d_agg_status, uint8_t* d_ok_out, void* stream);
/* Software-pipelined form of b2_epoch_dev.  Epoch k uses slot = k mod depth, depth <= B2_EPOCH_SLOTS:
 *   b2_epoch_start_dev(slot)  on `stream`: forks the pubkey/hash half, decompresses and segment-sums the signatures (waits until
 *                             the slot's previous tail has drained);
 *   b2_epoch_tail_dev(slot)   on the slot's own high-priority stream: inversion + compress, subgroup check, second Miller loop,
 *                             final exponentiation, update_latest_messages (applied in epoch order).  Its outputs (aggregate
 *                             signatures, verdicts, the LMD table) are valid for a stream only after b2_epoch_wait_dev(slot, it).
 *                             d_target_epoch == d_block_idx == NULL: no LMD update in the tail -- the form of a SHARDED epoch (one
 *                             validator set, committees split over ranks by slot, pos-evolution.md:455), whose caller all-gathers
 *                             the verdicts and then applies b2_latest_messages_update_dev for ALL aggregates on every rank;
 *   fork choice of epoch k    b2_epoch_wait_dev(slot, s); b2_vote_weights_dev(s); all-reduce; b2_head_from_votes_dev(s) -- enqueue it
 *                             before the tail of epoch k+1, whose LMD update waits for the vote scatter issued before it.
 * The latency-bound tails of epochs k, k-1, .. thereby overlap with each other and with the grid-filling signature decompression of
 * epoch k+1: under that contention one tail takes longer than one decompression, so depth 3 is what keeps the multiply pipe busy. */
#define B2_EPOCH_SLOTS 16
int b2_epoch_start_dev(b2_ctx* ctx, int slot, const uint8_t* d_sig96, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits,
                       uint32_t bits_stride, const uint8_t* d_msg32, uint32_t n_agg, uint64_t n_sig, int32_t* d_agg_status, void* stream);
int b2_epoch_tail_dev(b2_ctx* ctx, int slot, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride,
                      const uint64_t* d_target_epoch, const uint32_t* d_block_idx, uint32_t n_agg, uint8_t* d_agg_sig96, int32_t* d_agg_status,
                      uint8_t* d_ok_out);
int b2_epoch_wait_dev(b2_ctx* ctx, int slot, void* stream);
/* Which form of the pairing kernels (K5/K6) b2_epoch_start_dev / b2_epoch_tail_dev enqueue from now on: 0 = thread per aggregate
 * (fewest instructions: right while further epochs keep the chip busy), 1 = three lanes per pairing (shortest critical path: right for
 * the last epochs of a batch, whose tails drain with nothing left to overlap).  The synchronous entry points always use form 1. */
int b2_epoch_set_pairing_form(b2_ctx* ctx, int form);
int b2_latest_messages_update_dev(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits,
                                  uint32_t bits_stride, const uint64_t* d_target_epoch, const uint32_t* d_block_idx,
                                  const uint8_t* d_accept, uint32_t n_agg, void* stream);
/* direct (un-propagated) vote weight per block in tree pre-order, u64[n_blocks]: the quantity that is summed across
 * GPUs (ncclAllReduce sum, u64) when validators are sharded; then b2_head_from_votes_dev finishes on every rank. */
int b2_vote_weights_dev(b2_ctx* ctx, uint64_t* d_votes_preorder, void* stream);
/* the same for the validators [v_begin, v_end) only: the get_head shard of one rank when ONE validator set is spread over the GPUs of
 * a box (BASELINE.json configs 4/5: N/8 validators per GPU, u64[n_blocks] all-reduce, head on every rank; pos-evolution.md:1102-1116) */
int b2_vote_weights_range_dev(b2_ctx* ctx, uint64_t v_begin, uint64_t v_end, uint64_t* d_votes_preorder, void* stream);
int b2_head_from_votes_dev(b2_ctx* ctx, uint64_t* d_votes_preorder, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score,
                           uint64_t* d_weight_out /* may be NULL */, uint32_t* d_head_idx_out, void* stream);
/* ---- get_head over ONE validator set spread over the GPUs of a box (BASELINE.json config 4; pos-evolution.md:1102-1116), one kernel
 * per rank with the all-reduce of the vote weights fused in over NVLink peer memory (no NCCL launch between scatter and tree):
 *   b2_fc_exchange_export  allocates this rank's exchange block (accumulators + flags, sized by the loaded tree) and returns its
 *                          64-byte CUDA IPC handle; the host layer all-gathers the handles of the ranks (any transport);
 *   b2_fc_exchange_open    maps every peer's block (handles64 = world x 64 bytes, in rank order);
 *   b2_get_head_multi      COLLECTIVE (every rank, same order): scatter the votes of this rank's validators [v_begin, v_end),
 *                          push the per-block sums into every rank's accumulator with 64-bit reductions over NVLink, flag
 *                          exchange, tree phase on the complete accumulator; returns the same head on every rank.
 * One process per GPU, at most 8 ranks, all on one NVLink/NVSwitch domain.  Re-export after b2_tree_load. */
int b2_fc_exchange_export(b2_ctx* ctx, uint8_t* handle64_out);
int b2_fc_exchange_open(b2_ctx* ctx, int rank, int world, const uint8_t* handles64);
int b2_get_head_multi(b2_ctx* ctx, uint64_t v_begin, uint64_t v_end, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score,
                      uint32_t* head_idx_out);
uint32_t b2_tree_size(b2_ctx* ctx);
/* profiling aid: SM-clock stamps of the phases of the context's last get_head (vote scatter of CTA 0, tree phases); u64[32] */
int b2_debug_head_clocks(b2_ctx* ctx, uint64_t* out32);

#ifdef __cplusplus
}
#endif
#endif /* B200POS_H */
