/*
 * g2048.h — C ABI of the B200-native batched 2048 environment (libg2048.so).
 *
 * This is the drop-in boundary for ONE path of rgal/gym-2048: Game2048Env.step()
 * (+ the reset() it needs for auto-reset) batched over independent boards.  The
 * reference has no FFI layer of its own (it is pure Python); every entry point
 * below names the reference method it replaces as `env/envs/game2048_env.py:LINE`
 * and is what a ctypes/cffi binding in that file would call (INTEGRATION.md shows
 * the stub).
 *
 * Conventions
 *   - plain C: pointers, sizes, scalars.  No torch / CUDA types in any signature;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every pointer of the stateless API is CALLER-OWNED DEVICE memory on the
 *     current CUDA device; the library allocates nothing persistent there and
 *     launches asynchronously on `stream` (no implicit synchronisation).
 *   - return value: 0 on success, a negative G2048_ERR_* otherwise;
 *     g2048_last_error() gives the text (thread-local).  Nothing throws.
 *   - board encoding: 16 bytes per board, row-major cell (r,c) at byte 4*r+c,
 *     each byte the EXPONENT e of the tile value 2^e (0 = empty cell).  Board
 *     arrays must be 16-byte aligned.  g2048_values_from_exp/exp_from_values
 *     convert to/from the reference's 4x4 int64 tile VALUES (Matrix).
 *   - actions: 0=Up 1=Right 2=Down 3=Left (game2048_env.py:210-212).
 *   - determinism: spawn draws are counter-based Philox words addressed by
 *     (seed, global env id, step index); results do not depend on how the batch
 *     is sharded over GPUs (env_id_base = global id of element 0).
 *
 * Draw stream (frozen, version 2; ABI >= 3) — see DESIGN.md §3.  Philox as
 * published (Salmon et al., SC'11; Random123 constants and known-answer vectors):
 *   key      = philox4x32_10(ctr = {idx_hi, env_hi, tag, 0},
 *                            key = {seed_lo, seed_hi}).word0   [the same for a whole launch]
 *   (x0, x1) = philox2x32_10(ctr = {env_lo, idx_lo}, key)      [once per board]
 *   w[0] = x0,  w[1] = x1,  w[2] = x1 << 16,  w[3] = 0
 *   tag 0: idx = step_index (g2048_step, g2048_add_tile)
 *   tag 1: idx = reset_index (g2048_reset)
 *   tag 2: idx = step_index (g2048_sample_actions: policy word w[0])
 *   spawn(board, w): n = #empty cells; p = (uint64)w * n; k = p >> 32;
 *                    f = (uint32)p;  tile = (f < 3865470567u) ? 2 : 4   [P(2)=0.9]
 *                    the k-th empty cell in row-major order receives the tile.
 *   step spawn uses w[0]; a reset (auto-reset in g2048_step, or g2048_reset)
 *   zeroes the board and spawns with w[1] then w[2].
 *   (Version 1, ABI <= 2, drew all four words from one philox4x32_10 block per
 *   board; version 2 halves the generator's work per board.)
 */
#ifndef G2048_H
#define G2048_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2048_ABI_VERSION 5 /* 2: G2048StepArgs.boards_out, data-side entry points */
                            /* 3: draw stream version 2, g2048_philox2x32, g2048_draw_words, */
                            /*    g2048_step_many                                            */
                            /* 4: G2048StepArgs.ep_return/final_return, g2048_step_n,        */
                            /*    g2048_one (single-env packed call)                         */
                            /* 5: G2048StepArgs.chain, G2048_FLAG_CHAINED (chained launches)  */

#define G2048_TAG_STEP   0u
#define G2048_TAG_RESET  1u
#define G2048_TAG_POLICY 2u

#define G2048_OK            0
#define G2048_ERR_INVALID  (-1) /* bad argument (NULL required pointer, bad dtype, ...) */
#define G2048_ERR_ALIGN    (-2) /* board pointer not 16-byte aligned                   */
#define G2048_ERR_CUDA     (-3) /* CUDA runtime error (text in g2048_last_error)       */
#define G2048_ERR_NOMEM    (-4) /* allocation failed (stateful API only)               */

/* G2048StepArgs.flags */
#define G2048_FLAG_AUTO_RESET 1u /* SB3 DummyVecEnv semantics: a terminated env is   */
                                 /* replaced by a fresh reset() board in the same step */
/* g2048_step / _step_n / _step_list / _step_many: the kernel plays g2048_sample_actions'    */
/* policy itself and WRITES the action it drew (g2048_step: into `actions`)                */
#define G2048_FLAG_POLICY_UNIFORM 2u /* action uniform in {0,1,2,3} (train.py:119)        */
#define G2048_FLAG_POLICY_LEGAL   4u /* uniform among the legal moves of the live board   */
/*
 * Chained launches (G2048StepArgs.chain).  Consecutive steps of the SAME boards depend on each other board by
 * board; a kernel launch normally waits for the whole previous launch to drain (~2.3 us of an ~11 us step over
 * 1 Mi boards).  With a chain buffer every warp of the step kernel publishes "step_index + 1" in its own word of
 * `chain` when it is done, and a launch that carries G2048_FLAG_CHAINED waits warp by warp for the step
 * step_index - 1 instead of for the grid: its CTAs start as soon as an SM has room for them.  Results are
 * bit-identical to unchained launches (tests/test_gpu_chain.py).
 *   chain  device memory, G2048_CHAIN_BYTES, 8-byte aligned, ZEROED before its first use and whenever n changes;
 *          one buffer per env set (per `boards`), used on one stream.  Every launch that carries it — chained or
 *          not — publishes, so chained and plain launches of a set can be mixed freely.
 *   G2048_FLAG_CHAINED  the caller's promise about THIS launch: the previous launch that carried this chain
 *          buffer had the same n and step_index - 1, and everything this launch reads (boards, actions or — with
 *          G2048_FLAG_POLICY_LEGAL — legal_mask, the running episode statistics, forced_draws) was last written
 *          either by that launch or before it was issued.  Open-loop action sequences (g2048_step_n sets the
 *          flag itself from the second step on), the in-kernel policies, round-robin stepping of several env
 *          sets.  NOT a closed loop in which a policy kernel writes this step's actions after the previous step:
 *          leave the flag off there — the launch then waits for all earlier work on the stream as usual.  A
 *          launch with step_index 0 is never chained.  Not combinable with step_counter.
 *   G2048_FLAG_CHAIN_INTERLEAVED  how the chain is used, which decides the launch shape (a chained launch must
 *          carry the same value as the launch it chains to; change it with an unchained launch).  Clear: the
 *          launches of this chain follow each other directly on the stream (g2048_step_n, one env set stepped
 *          in a loop) — every launch fills the machine.  Set: launches of OTHER chains sit in between (several
 *          env sets stepped round-robin) — a launch takes a fifth of every SM and up to five consecutive launches
 *          run side by side (it pays from two alternating sets on, profiles/r02_chain_sets.log).  A wrong choice
 *          costs time, not correctness.
 */
#define G2048_FLAG_CHAINED 8u
#define G2048_FLAG_CHAIN_INTERLEAVED 16u
#define G2048_CHAIN_WORDS 16384u                   /* 64-bit words */
#define G2048_CHAIN_BYTES (8u * G2048_CHAIN_WORDS)

/* g2048_encode_obs dtype */
#define G2048_OBS_U8   0
#define G2048_OBS_F32  1
#define G2048_OBS_I64  2 /* the reference's dtype (numpy int on Linux) */
#define G2048_OBS_BF16 3

#define G2048_P2_THRESHOLD 3865470567u /* f < this  <=>  f / 2^32 < 0.9  */

/*
 * One batched step.  Replaces Game2048Env.step (game2048_env.py:76-100) incl.
 * move (:194-241), shift (:243-260), add_tile (:166-176), isend (:262-280),
 * highest (:190-192) and — with G2048_FLAG_AUTO_RESET — reset (:102-111).
 */
typedef struct G2048StepArgs {
  uint8_t*        boards;          /* [n*16] in/out (in only when boards_out is set)       */
  const uint8_t*  actions;         /* [n]    0..3; only the low 2 bits are read.  With a   */
                                   /*        G2048_FLAG_POLICY_* flag this array is an     */
                                   /*        OUTPUT: the kernel draws the action exactly as */
                                   /*        g2048_sample_actions(legal_mask or NULL, ...,  */
                                   /*        step_index) would and stores it here; _LEGAL   */
                                   /*        needs legal_mask, which then is in (the mask   */
                                   /*        of the board being stepped) AND out (the mask  */
                                   /*        of the board handed back): BASELINE config 4   */
                                   /*        as one launch per step                         */
  float*          rewards;         /* [n]    out: merge score, or illegal_move_reward      */
  uint8_t*        dones;           /* [n]    out: terminated (0/1)                         */
  uint8_t*        illegal;         /* [n]    out, nullable: info['illegal_move']           */
  uint8_t*        highest_exp;     /* [n]    out, nullable: log2(info['highest']) of the   */
                                   /*        post-spawn (terminal, pre-reset) board        */
  uint8_t*        legal_mask;      /* [n]    out, nullable: bit d = move d legal on the    */
                                   /*        board RETURNED in `boards`                    */
  uint8_t*        terminal_boards; /* [n*16] out, nullable: written only where done        */
  uint32_t*       ep_score;        /* [n]    in/out, nullable: running episode score       */
  uint32_t*       ep_len;          /* [n]    in/out, nullable: running episode length      */
  uint32_t*       final_score;     /* [n]    out, nullable: episode score, where done      */
  uint32_t*       final_len;       /* [n]    out, nullable: episode length, where done     */
  const uint32_t* forced_draws;    /* [n*4]  nullable: words used INSTEAD of the draw      */
                                   /*        words w[0..3] (fixture / CSV parity)          */
  uint64_t*       step_counter;    /* nullable device uint64[2]: when set the step index is */
                                   /*        read from step_counter[0] instead of step_index */
                                   /*        and the launch itself advances it when it ends  */
                                   /*        (CUDA-graph replay); step_counter[1] is scratch */
                                   /*        of the library and must be 0 before the 1st use */
  uint64_t        n;               /* boards in this call                                  */
  uint64_t        env_id_base;     /* global env id of element 0                           */
  uint64_t        seed;            /* draw-stream seed                                     */
  uint64_t        step_index;      /* draw-stream index; caller increments                 */
  float           illegal_move_reward; /* set_illegal_move_reward (:61-67), default 0      */
  uint32_t        max_tile_exp;    /* set_max_tile (:69-73) as exponent; 0 = None          */
  uint32_t        flags;           /* G2048_FLAG_*                                         */
  uint8_t*        boards_out;      /* [n*16] out, nullable: where the boards handed back   */
                                   /*        to the agent are written; NULL = in place.    */
                                   /*        Lets a rollout/trajectory buffer slice t+1 be */
                                   /*        produced from slice t with no copy.           */
  float*          ep_return;       /* [n]    in/out, nullable: running sum of the REWARDS  */
                                   /*        the step emitted (illegal_move_reward included */
                                   /*        — what SB3's Monitor sums, ppo_train.py:123)   */
  float*          final_return;    /* [n]    out, nullable: that sum, where done           */
  uint8_t*        boards_nibble;   /* [n*8]  out, nullable: the board handed back, packed  */
                                   /*        4 bits per cell — cell c = bits 4c..4c+3 of   */
                                   /*        the little-endian 64-bit word, the classic    */
                                   /*        2048 bitboard; holds exponents <= 15 (tiles   */
                                   /*        <= 32768); 8-byte aligned                     */
  uint32_t*       nibble_overflow; /* device uint32, nullable: incremented once per board  */
                                   /*        whose exponents do not fit (a tile >= 65536)  */
  uint64_t*       chain;           /* device, G2048_CHAIN_BYTES, nullable: chained launches */
                                   /*        (G2048_FLAG_CHAINED above)                     */
} G2048StepArgs;

int g2048_abi_version(void);
const char* g2048_last_error(void);

int g2048_step(const G2048StepArgs* args, void* stream);

/*
 * n_steps consecutive steps of the same boards, ONE KERNEL LAUNCH PER STEP (each step reads
 * and writes the boards in device memory, 38 algorithmic bytes per board like g2048_step),
 * issued back to back from C as programmatic dependent launches.  Exactly what this loop does
 * — which is the loop SB3's DummyVecEnv runs behind ppo_train.py:123, minus the interpreter
 * between two steps:
 *     for (k = 0; k < n_steps; ++k) { g2048_step(&a, stream); advance(a); }
 * where advance() moves every PER-STEP array on by row_stride elements (actions, rewards,
 * dones, illegal, highest_exp, legal_mask, final_score, final_len, final_return; boards for
 * terminal_boards; 4 words for forced_draws) and step_index by 1; the per-env state (boards,
 * ep_score, ep_len, ep_return) stays.  row_stride = 0 reuses the same rows every step;
 * otherwise row_stride >= n.  boards_out and step_counter must be NULL.
 * Use: open-loop action sequences stepped at the launch rate of C instead of the caller's
 * language (a 131,072-board shard is stepped in ~3 us; a Python loop issues a launch every
 * ~6 us).  g2048_step_many is the variant that also keeps the boards in registers.
 * With args->chain set the steps are chained launches (G2048_FLAG_CHAINED above): the first step carries the
 * caller's flags, every later one the flag as well — step t+1 reads what step t wrote and rows that were there
 * before the call.
 */
int g2048_step_n(const G2048StepArgs* args, uint32_t n_steps, uint64_t row_stride, void* stream);

/*
 * A caller-built list of `count` complete g2048_step calls (each element with its own
 * boards, per-step rows, env_id_base, step_index ...), launched in order on `stream` by ONE
 * call — the general form of g2048_step_n: the elements may step different env sets (several
 * vectorised envs stepped round-robin) or the same one (then element j+1 must carry
 * step_index + 1).  Fails on the first invalid element; earlier elements stay launched.
 */
int g2048_step_list(const G2048StepArgs* list, uint64_t count, void* stream);
/*
 * The same, with CUDA events (cudaEvent_t handles as void*) recorded on the stream: events[0]
 * before the first launch, events[r] after launch r*every (r = 1 .. n_events-1, as far as the
 * list goes).  A benchmark's R timed regions of `every` launches are one C call: on shards a
 * GPU steps in ~3 us the interpreter's ~10 us per region would otherwise be what is measured.
 */
int g2048_step_list_timed(const G2048StepArgs* list, uint64_t count, uint64_t every,
                          void* const* events, uint64_t n_events, void* stream);

/*
 * n_steps steps in one launch, for open-loop action sequences (pre-generated random
 * actions as in `train.py:119`, replays of recorded games): exactly what n_steps calls of
 * g2048_step with actions + k*n, rewards + k*n, dones + k*n and step_index + k would
 * leave — same draws, same auto-reset — but the boards stay in registers between
 * the steps, so a step costs one action byte read and a reward and a done written
 * instead of 38 bytes, and there is one launch instead of n_steps.  The per-step
 * arrays are step-major: element (k, i) at k*n + i.
 * With G2048_FLAG_POLICY_UNIFORM / _LEGAL the actions are not read but drawn in the
 * kernel, exactly as g2048_sample_actions(legal mask of the live board, ..., step_index + k)
 * would draw them before step k (policy-tag stream), and written to actions_out: a whole
 * random rollout — `train.py:119`'s random agent, BASELINE config 4's random-legal play —
 * is one launch.
 */
typedef struct G2048StepManyArgs {
  uint8_t*       boards;       /* [n*16]         in/out: before the first / after the last step */
  const uint8_t* actions;      /* [n_steps*n]    0..3; only the low 2 bits are read; NULL with  */
                               /*                a policy flag                                  */
  float*         rewards;      /* [n_steps*n]    out                                            */
  uint8_t*       dones;        /* [n_steps*n]    out                                            */
  uint8_t*       illegal;      /* [n_steps*n]    out, nullable                                  */
  uint8_t*       boards_traj;  /* [n_steps*n*16] out, nullable: the board handed back to the    */
                               /*                agent after every step (row k = after step k)  */
  uint64_t       n;
  uint64_t       env_id_base;
  uint64_t       seed;
  uint64_t       step_index;   /* index of the first step; step k uses step_index + k           */
  uint32_t       n_steps;
  float          illegal_move_reward;
  uint32_t       max_tile_exp;
  uint32_t       flags;        /* G2048_FLAG_*                                                  */
  uint8_t*       actions_out;  /* [n_steps*n]    out, nullable: the actions a policy flag drew  */
  uint8_t*       legal_mask;   /* [n_steps*n]    out, nullable: bit d = move d legal on the     */
                               /*                board handed back after the step               */
} G2048StepManyArgs;

int g2048_step_many(const G2048StepManyArgs* args, void* stream);

/*
 * Game2048Env.reset (game2048_env.py:102-111) for every env whose reset_mask
 * byte is non-zero (all envs when reset_mask is NULL): zero board + two spawns.
 */
int g2048_reset(uint8_t* boards, const uint8_t* reset_mask, uint64_t n,
                uint64_t env_id_base, uint64_t seed, uint64_t reset_index,
                void* stream);

/*
 * Game2048Env.add_tile (game2048_env.py:166-176) on its own: one spawn per board with the
 * step-tag draw word w[0] at `step_index` (boards without an empty cell are left alone,
 * where the reference asserts, :176).
 */
int g2048_add_tile(uint8_t* boards, uint64_t n, uint64_t env_id_base, uint64_t seed,
                   uint64_t step_index, void* stream);

/*
 * Game2048Env.move(direction, trial) (game2048_env.py:194-241) without spawn:
 * boards_out may alias boards_in (trial=False) or be NULL (trial=True).
 * changed[i]==0 is the reference's IllegalMove.  scores/changed nullable.
 */
int g2048_move(const uint8_t* boards_in, uint8_t* boards_out,
               const uint8_t* directions, uint32_t* scores, uint8_t* changed,
               uint64_t n, void* stream);

/*
 * Board queries: legal-move mask (move(d, trial=True) for d=0..3, :224,236-239),
 * highest (:190-192) as exponent, number of empties (:186-188), isend (:262-280).
 * Every output is nullable.
 */
int g2048_status(const uint8_t* boards, uint8_t* legal_mask, uint8_t* highest_exp,
                 uint8_t* n_empty, uint8_t* is_end, uint32_t max_tile_exp,
                 uint64_t n, void* stream);

/*
 * stack(flat, layers=15) (game2048_env.py:17-32): one-hot [n,16,4,4], channel 0 =
 * empty, channel k = (cell == 2^k), k=1..15; a tile >= 2^16 lights no channel.
 */
int g2048_encode_obs(const uint8_t* boards, void* obs, int dtype, uint64_t n,
                     void* stream);

/*
 * Tile values (reference Matrix, int64) <-> exponents.  exp_from_values counts
 * cells that are neither 0 nor a power of two 2..2^31 into *bad_count (device
 * uint32, nullable) and stores 0 for them.
 */
int g2048_values_from_exp(const uint8_t* boards, int64_t* values, uint64_t n_cells,
                          void* stream);
int g2048_exp_from_values(const int64_t* values, uint8_t* boards, uint64_t n_cells,
                          uint32_t* bad_count, void* stream);

/*
 * ONE env per call: the reference's single-env class (game2048_env.py:34-288) on a packed
 * in/out block that may live in pinned HOST memory (cudaHostAlloc: the device reads and
 * writes it in place, zero copy), so a call costs one launch and — with synchronize != 0 —
 * one stream synchronisation.  `values` is the reference's Matrix (tile VALUES, row-major).
 *   G2048_ONE_STEP      step(action) (:76-100) without auto-reset: move, spawn with the step-tag
 *                       draw word of (seed, env 0, index), isend; an illegal move leaves values
 *                       untouched, sets illegal/done and reward = illegal_move_reward
 *   G2048_ONE_RESET     reset() (:102-111): values = fresh board from the reset-tag words at `index`
 *   G2048_ONE_MOVE      move(action, trial) (:194-241): no spawn; changed == 0 is IllegalMove;
 *                       values are written only when changed and trial == 0; reward = score
 *   G2048_ONE_ADD_TILE  add_tile() (:166-176) with the step-tag word at `index`
 *   G2048_ONE_STATUS    nothing moves: only the queries below
 * Every op also fills highest_exp (:190-192), n_empty (:186-188), is_end (:262-280),
 * legal_mask and obs = stack(values) (:17-32, int64 [16][4][4]) of the values as they are
 * after the op.  Cells that are not 0 or a power of two 2..2^31 are counted in bad_cells; a
 * board with such cells is not stepped/moved (obs still follows the reference: they light no
 * channel).
 */
#define G2048_ONE_STEP     0
#define G2048_ONE_RESET    1
#define G2048_ONE_MOVE     2
#define G2048_ONE_ADD_TILE 3
#define G2048_ONE_STATUS   4
typedef struct G2048OneIO {
  int64_t  values[16];   /* in/out */
  int64_t  obs[256];     /* out    */
  float    reward;       /* out: STEP merge score or illegal_move_reward; MOVE score */
  uint32_t score;        /* out: merge score of the move                             */
  uint8_t  done;         /* out: STEP terminated                                     */
  uint8_t  illegal;      /* out: STEP info['illegal_move']                           */
  uint8_t  changed;      /* out: the op changed the board                            */
  uint8_t  highest_exp;  /* out */
  uint8_t  legal_mask;   /* out */
  uint8_t  n_empty;      /* out */
  uint8_t  is_end;       /* out */
  uint8_t  bad_cells;    /* out */
} G2048OneIO;
int g2048_one(G2048OneIO* io, int op, int action, int trial, uint64_t seed, uint64_t index,
              float illegal_move_reward, uint32_t max_tile_exp, void* stream, int synchronize);

/* Debug: out[4*i..4*i+3] = philox4x32_10(ctr[4*i..], key) — known-answer tests. */
int g2048_philox(const uint32_t* ctr, uint32_t key0, uint32_t key1, uint32_t* out,
                 uint64_t n, void* stream);
/* Debug: out[2*i..2*i+1] = philox2x32_10(ctr[2*i..], key) — known-answer tests (8-byte aligned). */
int g2048_philox2x32(const uint32_t* ctr, uint32_t key, uint32_t* out, uint64_t n,
                     void* stream);
/* The draw words w[0..3] of env ids env_id_base .. env_id_base+n-1 at `index` under `tag`
 * (G2048_TAG_*), as the kernels compute them: words[4*i..4*i+3], 16-byte aligned. */
int g2048_draw_words(uint32_t* words, uint64_t n, uint64_t env_id_base, uint64_t seed,
                     uint64_t index, uint32_t tag, void* stream);

/* ------------------------------------------------------------------------- */
/* Either side of the step: the random policies that drive it and the        */
/* transition data it produces (reference: train.py, training_data.py).      */
/* ------------------------------------------------------------------------- */

/*
 * Uniform-random policy on the device.  legal_mask NULL: action uniform in
 * {0,1,2,3} (train.py:119 `random.randint(0, 3)`, the benchmark's random
 * actions).  legal_mask given: uniform among the set bits of legal_mask[i] (all
 * four when the mask is 0) — the random-legal policy of BASELINE config 4.
 * action = the k-th set bit, k = hi32(w[0] * popcount), w = the policy-tag draw
 * words of (seed, env id, step_index): a stream of its own, independent of the
 * spawn g2048_step makes at the same index.
 */
int g2048_sample_actions(const uint8_t* legal_mask, uint8_t* actions, uint64_t n,
                         uint64_t env_id_base, uint64_t seed, uint64_t step_index,
                         void* stream);

/*
 * One board symmetry applied to n transitions: training_data.hflip
 * (training_data.py:257-272: columns reversed, actions 1<->3) when hflip != 0,
 * then training_data.rotate(k) (:274-279: np.rot90(k, axes=(2,1)), action + k
 * mod 4).  next_in/next_out and actions_in/actions_out are nullable pairs; out
 * may alias in.
 */
int g2048_symmetry(const uint8_t* boards_in, uint8_t* boards_out,
                   const uint8_t* next_in, uint8_t* next_out,
                   const uint8_t* actions_in, uint8_t* actions_out, uint64_t n,
                   int hflip, int k, void* stream);

/*
 * training_data.augment (training_data.py:281-299): the 8 symmetric copies of n
 * transitions, laid out as the reference's merge order [X, H, R1 X, R1 H, R2 X,
 * R2 H, R3 X, R3 H] (H = hflip, Rk = rotate(k)): copy s = 2k+h occupies rows
 * [s*n, (s+1)*n) of every output ([8n*16] boards, [8n] actions/rewards/dones).
 */
int g2048_augment(const uint8_t* boards, const uint8_t* next_boards,
                  const uint8_t* actions, const float* rewards, const uint8_t* dones,
                  uint64_t n, uint8_t* boards_out, uint8_t* next_boards_out,
                  uint8_t* actions_out, float* rewards_out, uint8_t* dones_out,
                  void* stream);

/*
 * training_data.get_discounted_return (training_data.py:104-124) over n rows in
 * game order: G[i] = r[i] if dones[i] or i == n-1, else r[i] + gamma * G[i+1],
 * evaluated in float64 in the reference's order (no fused multiply-add), one
 * thread per episode segment.
 */
int g2048_discounted_return(const float* rewards, const uint8_t* dones,
                            double* returns, uint64_t n, double gamma, void* stream);

/*
 * Generalised advantage estimation over a [T,n] rollout (time-major, env fastest), the
 * computation SB3's RolloutBuffer.compute_returns_and_advantage runs after the rollout of
 * ppo_train.py:138-183 (third-party; restated from its published algorithm):
 *   for t = T-1..0: nn = 1 - (t == T-1 ? last_dones : episode_starts[t+1]);
 *                   nv = (t == T-1 ? last_values : values[t+1]);
 *                   delta = r[t] + gamma*nv*nn - v[t];  A[t] = delta + gamma*lambda*nn*A[t+1];
 *   returns = A + values.   float32 with gamma and gamma*lambda (product in double) rounded
 *   to float32 first, each product/sum rounded separately (no fma): the result equals the
 *   numpy float32 expression evaluated op by op.  One thread per env.
 */
int g2048_gae(const float* rewards, const float* values, const uint8_t* episode_starts,
              const float* last_values, const uint8_t* last_dones, float* advantages,
              float* returns, uint64_t T, uint64_t n, double gamma, double gae_lambda,
              void* stream);

/*
 * The reference's transition CSV (training_data.export_csv / import_csv,
 * training_data.py:188-248), HOST buffers: 16 board tile VALUES, action, reward
 * ("%f"), 16 next-board values, done, optionally the discounted return; header
 * "1-1,...,4-4,action,reward,next 1-1,...,next 4-4,done[,return]".  Boards are
 * exponents on our side and tile values in the file.  export writes the header
 * unless append != 0.  rows() counts data rows; import fills caller buffers of
 * n_rows entries (returns nullable) and fails on cells that are not 0 or a
 * power of two.
 */
int g2048_csv_export(const char* path, const uint8_t* boards, const uint8_t* actions,
                     const double* rewards, const uint8_t* next_boards,
                     const uint8_t* dones, const double* returns, uint64_t n,
                     int append);
int g2048_csv_rows(const char* path, uint64_t* n_rows, int* has_returns);
int g2048_csv_import(const char* path, uint8_t* boards, uint8_t* actions,
                     double* rewards, uint8_t* next_boards, uint8_t* dones,
                     double* returns, uint64_t n_rows);

/* ------------------------------------------------------------------------- */
/* Stateful host-buffer API: what a non-CUDA host (the reference's Python, a  */
/* cgo/JNI caller) binds.  The handle owns the device state for n boards plus */
/* pinned staging; actions come from and results go to HOST memory.           */
/* ------------------------------------------------------------------------- */
typedef struct G2048Env G2048Env;

typedef struct G2048EnvConfig {
  int32_t  device;               /* CUDA device ordinal                              */
  uint32_t flags;                /* G2048_FLAG_*                                     */
  uint64_t n;                    /* boards owned by this handle                      */
  uint64_t env_id_base;          /* global id of board 0 (sharding)                  */
  uint64_t seed;
  float    illegal_move_reward;
  uint32_t max_tile_exp;
  uint32_t n_chunks;             /* copy/compute pipeline depth; 0 = library default (2: a 1/16 lead slice + the rest) */
  uint32_t board_format;         /* G2048_BOARDS_BYTES (0), _NIBBLE or _BYTES_PACKED_WIRE: what                */
                                 /* g2048_env_step_host writes to out->boards and how it gets there            */
  uint32_t unpack_threads;       /* G2048_BOARDS_BYTES_PACKED_WIRE: host threads that expand the boards;       */
                                 /* 0 = library default (half the CPUs the process may run on, at most 8;      */
                                 /* all but one of them on hosts with 8 CPUs or fewer)                         */
} G2048EnvConfig;

/* G2048EnvConfig.board_format.  The step is PCIe-bound for a host caller (21 bytes per board come back);  */
/* nibble boards cut that to 13: out->boards is then [n*8], 4 bits per cell (see G2048StepArgs.boards_     */
/* nibble).  A board holding a tile >= 65536 does not fit: such boards are counted in                     */
/* out->nibble_overflow and g2048_env_get_boards_host returns the full 16-byte boards.                    */
#define G2048_BOARDS_BYTES  0u
#define G2048_BOARDS_NIBBLE 1u
/* out->boards is [n*16] exponent bytes exactly as with G2048_BOARDS_BYTES, but the boards cross PCIe 4 bits per cell   */
/* (13 instead of 21 bytes per board) and the library expands them into the caller's buffer with its own host threads,  */
/* slice by slice, while the next slice is still on the wire.  A step in which some board holds a tile >= 65536 is      */
/* detected (the kernel counts them) and its boards are fetched again in full: correct, at the plain format's speed.    */
#define G2048_BOARDS_BYTES_PACKED_WIRE 2u

/* Host result pointers for g2048_env_step_host; boards/rewards/dones required. */
typedef struct G2048HostStepOut {
  uint8_t*  boards;       /* [n*16], or [n*8] with G2048_BOARDS_NIBBLE */
  float*    rewards;      /* [n]    */
  uint8_t*  dones;        /* [n]    */
  uint8_t*  illegal;      /* [n] nullable */
  uint8_t*  highest_exp;  /* [n] nullable */
  uint8_t*  legal_mask;   /* [n] nullable */
  uint32_t* nibble_overflow; /* host uint32, nullable: boards of this step that do not fit G2048_BOARDS_NIBBLE */
} G2048HostStepOut;

int g2048_env_create(G2048Env** out, const G2048EnvConfig* cfg);
int g2048_env_destroy(G2048Env* env);
/* reset all boards (advances the handle's reset counter); boards_host nullable */
int g2048_env_reset_host(G2048Env* env, uint8_t* boards_host);
/* one step: H2D actions -> kernel -> D2H results, chunk-pipelined; synchronous */
int g2048_env_step_host(G2048Env* env, const uint8_t* actions_host,
                        const G2048HostStepOut* out);
/* device pointers of the handle's state (for zero-copy hosts such as torch)   */
int g2048_env_device_ptrs(G2048Env* env, uint8_t** boards, float** rewards,
                          uint8_t** dones);
int g2048_env_set_boards_host(G2048Env* env, const uint8_t* boards_host);
/* the live boards in the 16-byte format, whatever board_format is (synchronous) */
int g2048_env_get_boards_host(G2048Env* env, uint8_t* boards_host);
uint64_t g2048_env_step_index(const G2048Env* env);

/* Host-only helper (no device involved): expand n boards from 4 bits per cell ([n*8], the layout of                 */
/* G2048StepArgs.boards_nibble / G2048_BOARDS_NIBBLE) to one exponent byte per cell ([n*16]) — what the library's   */
/* own threads do for G2048_BOARDS_BYTES_PACKED_WIRE, for a caller that took the packed boards.                     */
int g2048_unpack_boards_host(const uint8_t* packed, uint8_t* boards, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* G2048_H */
