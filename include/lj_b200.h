/*
 * lj_b200.h -- C ABI of the B200-native Lennard-Jones force + neighbour-list path.
 *
 * This is the drop-in boundary for the ONE hot path of kohnakagawa/lj_gpu: the kernel launch
 * inside measure() (cuda/force_cuda.cu:334) and the neighbour-list build that feeds it
 * (cuda/force_cuda.cu:122-163).  Every entry point names the reference interface it replaces.
 * Plain pointers and sizes only; all device work is asynchronous on the caller's stream
 * (a cudaStream_t passed as void*, NULL = the CUDA legacy default stream).  No function exits the
 * process: each returns an lj_status and leaves a message in lj_last_error_string().
 *
 * Ownership: the caller owns every array it passes (as the reference driver owns its
 * cuda_ptr globals, cuda/force_cuda.cu:24-30).  The library owns its context, streams,
 * memory pool and scratch (cell tables, staging ring).  It keeps no hidden copy of p.
 */
#ifndef LJ_B200_H
#define LJ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LJ_API __attribute__((visibility("default")))

typedef struct lj_ctx lj_ctx;

/* reference: checkCudaErrors -> message + exit(EXIT_FAILURE) (cuda/cuda_ptr.cuh:39-73);
 * here: a status code, the driver decides whether to exit. */
typedef enum lj_status {
  LJ_OK = 0,
  LJ_ERR_CUDA = 1,         /* a CUDA runtime call or kernel launch failed            */
  LJ_ERR_BAD_ARG = 2,      /* null pointer, negative size, unknown enum, misalignment */
  LJ_ERR_CAPACITY = 3,     /* sorted_list / ELL capacity too small (nothing written OOB) */
  LJ_ERR_OVERFLOW32 = 4,   /* list offsets do not fit the int32 pointer[] requested   */
  LJ_ERR_NO_DEVICE = 5,    /* no CUDA device: there is NO CPU fallback                */
  LJ_ERR_INVALID_LIST = 6  /* lj_validate_list found an out-of-range entry            */
} lj_status;

/* Particle-vector layout of q and p.  Reference: template parameter Vec of every kernel
 * (cuda/kernel.cuh:5), instantiated for double3 and double4 (cuda/force_cuda.cu:355-375);
 * SoA is the cpu_ref / OpenACC layout (cpu_ref/force_soa.cpp:15-18,
 * openacc/force_oacc_soa.cpp:17-22); float4 buffers exist but are never timed
 * (cuda/force_cuda.cu:24-25). */
typedef enum lj_layout {
  LJ_AOS_D3 = 0, /* packed {x,y,z} doubles, 24 B stride                                   */
  LJ_AOS_D4 = 1, /* {x,y,z,w} doubles, 32 B stride, 32 B aligned; .w ignored on read and
                    preserved on the write of p                                          */
  LJ_SOA_D = 2,  /* planes x[], y[], z[] of doubles: base + c*plane_stride + i            */
  LJ_AOS_F4 = 3, /* {x,y,z,w} floats, 16 B stride: list build, and force with LJ_PREC_MIXED
                    (FP32 pair math, p accumulated in float, one rounding per step)      */
  LJ_AOS_F3 = 4  /* packed {x,y,z} floats, 12 B stride (q_f3 / p_f3, cuda/force_cuda.cu:24):
                    list build, and force with LJ_PREC_MIXED, like LJ_AOS_F4             */
} lj_layout;

/* Neighbour-list storage.  CSR = sorted_list + number_of_partners + pointer (no sentinel,
 * cuda/force_cuda.cu:146-162); ELL = column-major transposed_list[i + k*pn], zero padded
 * (cuda/force_cuda.cu:229-240), pointer unused (the reference passes nullptr, :430-436). */
/* ELL_ROWS = row-major padded table list[i*ell_width + k], zero padded, the layout
 * make_sorted_list2d() aims at (cuda/force_cuda.cu:242-253) -- there with a fixed width of 60 that
 * is smaller than the longest row (rows overlap; no reference kernel reads it).  Here the width
 * is checked against max_partners by lj_build_ell_rows(). */
typedef enum lj_list_layout { LJ_LIST_CSR = 0, LJ_LIST_ELL = 1, LJ_LIST_ELL_ROWS = 2 } lj_list_layout;

/* Thread mapping.  Replaces the choice among the 20 kernels of cuda/kernel.cuh. */
typedef enum lj_variant {
  LJ_VARIANT_AUTO = 0,      /* library picks per layout/list/precision                      */
  LJ_VARIANT_SUBWARP = 1,   /* `group` lanes per i-particle (1,2,4,8,16,32); 32 = warp-per-i
                               (kernel.cuh:821-904), 1 = thread-per-i (kernel.cuh:67-236)     */
  LJ_VARIANT_TILE_TMA = 2,  /* CTA tile of rows, j-indices staged in shared memory by a TMA
                               bulk copy, `group` lanes per i (CSR only)                     */
  LJ_VARIANT_NEWTON3 = 3,   /* half list, reaction scattered with FP64 atomics
                               (the *_with_aar kernels, kernel.cuh:238-469): CSR with `group`
                               lanes per i, or the half ELL table with one thread per i
                               (memopt2/memopt3_with_aar, kernel.cuh:344-423)                */
  LJ_VARIANT_CLUSTER = 4,   /* cluster pair list built by lj_build_list(LJ_LIST_CLUSTERS) for
                               exactly these list arrays; error if there is none             */
  LJ_VARIANT_CELLTILE = 5   /* cell-tile mirror built by lj_build_list(LJ_LIST_TILES) for exactly
                               these list arrays: q[j] of a tile's neighbourhood staged in shared
                               memory by TMA, 16-bit local indices; error if there is none.
                               AUTO picks it whenever the mirror exists (FP64 and mixed)      */
} lj_variant;

typedef enum lj_precision {
  LJ_PREC_FP64 = 0,  /* all arithmetic FP64 (the reference's Dtype = double)                */
  LJ_PREC_MIXED = 1  /* FP32 pair arithmetic on 32-bit fixed-point coordinates (exact differences),
                        FP64 accumulation of the momenta; cutoff decisions near r2 == cl2 in FP64 */
} lj_precision;

/* ---------------------------------------------------------------- context ------------- */
/* Creates a context on `device` (a CUDA ordinal).  LJ_ERR_NO_DEVICE when CUDA is absent. */
LJ_API int lj_ctx_create(lj_ctx** out, int device);
LJ_API int lj_ctx_destroy(lj_ctx* ctx);
/* replaces cudaDeviceSynchronize() in measure() (cuda/force_cuda.cu:336); stream NULL = the default stream and
 * every stream the context owns */
LJ_API int lj_sync(lj_ctx* ctx, void* stream);
LJ_API const char* lj_last_error_string(lj_ctx* ctx);
LJ_API const char* lj_status_string(int status);
/* number of kernels this library has launched through `ctx` since creation */
LJ_API int64_t lj_launch_count(lj_ctx* ctx);
/* Live timing of the DOMINANT force kernel alone (lj_celltile_force, the kernel AUTO runs on large systems):
 * while enabled, every launch of it outside a stream capture is bracketed by two CUDA events on the launching
 * stream.  lj_kernel_timing(ctx, 1) starts with empty sums, (ctx, 0) stops; lj_kernel_timing_read() waits for
 * the recorded launches and returns their summed duration and their number.  Measurement aid for the roofline
 * of bench.py (the reference has no counterpart: its measure() times the whole loop, cuda/force_cuda.cu:319-342). */
LJ_API int lj_kernel_timing(lj_ctx* ctx, int enable);
LJ_API int lj_kernel_timing_read(lj_ctx* ctx, double* total_ms_out, int64_t* launches_out);
/* the context's own non-blocking stream (a cudaStream_t) */
LJ_API void* lj_ctx_stream(lj_ctx* ctx);
LJ_API int lj_device_sm_count(lj_ctx* ctx);

/* ---------------------------------------------------------------- memory layer -------- */
/* Replaces cuda_ptr<T> (cuda/cuda_ptr.cuh:10-104): a device allocation paired with a
 * pinned host mirror.  Device memory comes from a stream-ordered pool (cudaMallocAsync)
 * sized by the request, not by a static maximum. */
typedef struct lj_buf {
  void* host;   /* pinned (cuda_ptr::host_ptr, operator[])  */
  void* dev;    /* device (cuda_ptr::dev_ptr, operator T*)  */
  size_t bytes;
} lj_buf;

LJ_API int lj_buf_allocate(lj_ctx* ctx, size_t bytes, lj_buf* out, void* stream);   /* allocate()   */
LJ_API int lj_buf_deallocate(lj_ctx* ctx, lj_buf* buf, void* stream);               /* deallocate() */
/* host2dev(beg,count) / host2dev_async: byte range [beg, beg+count) */
LJ_API int lj_buf_host2dev(lj_ctx* ctx, const lj_buf* buf, size_t beg, size_t count, void* stream);
LJ_API int lj_buf_dev2host(lj_ctx* ctx, const lj_buf* buf, size_t beg, size_t count, void* stream);
/* set_val(beg,count,val) for 4-byte elements: fills host mirror and device range */
LJ_API int lj_buf_set_val32(lj_ctx* ctx, const lj_buf* buf, size_t beg_elems, size_t count_elems,
                     uint32_t value, void* stream);
/* raw stream-ordered device allocation for callers that keep their own host arrays */
LJ_API int lj_dev_alloc(lj_ctx* ctx, size_t bytes, void** out, void* stream);
LJ_API int lj_dev_free(lj_ctx* ctx, void* ptr, void* stream);
/* pageable host memory <-> device through the context's pinned double-buffered staging ring
 * (chunked, copy of chunk k+1 into the ring overlaps the DMA of chunk k) */
LJ_API int lj_upload(lj_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes, void* stream);
LJ_API int lj_download(lj_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes, void* stream);

/* ---------------------------------------------------------------- force step ---------- */
/* Replaces the launch
 *   kernel<<<block_num, THREAD_BLOCK>>>(q, p, particle_number, dt, CL2, list,
 *                                       number_of_partners, partner_pointer)
 * (cuda/force_cuda.cu:334; kernel signature cuda/kernel.cuh:5-13).  In place:
 *   p[i] += dt * sum_k f(r_ik) * (q[j_k] - q[i]),  f = (24 r^6 - 48)/r^14, masked to 0
 *   where r^2 > cl2.
 * q is read-only.  Rows may be in any order (the reference shuffles them,
 * cuda/force_cuda.cu:255-263).  All pointers are DEVICE pointers. */
typedef struct lj_force_args {
  const void* q;
  void* p;
  int64_t pn;                        /* particle_number                                    */
  double dt;
  double cl2;                        /* CL2 = cutoff^2                                     */
  const int32_t* list;               /* sorted_list (CSR) or transposed_list (ELL)         */
  const int32_t* number_of_partners;
  const void* pointer;               /* int32[pn] or int64[pn] (pointer64); NULL for ELL / ELL_ROWS */
  int32_t layout;                    /* lj_layout of q and p                               */
  int32_t list_layout;               /* lj_list_layout                                     */
  int32_t variant;                   /* lj_variant                                         */
  int32_t group;                     /* lanes per i-particle, 0 = default for the variant  */
  int32_t precision;                 /* lj_precision                                       */
  int32_t pointer64;                 /* 0: pointer is int32[], 1: int64[]                  */
  int32_t threads_per_block;         /* THREAD_BLOCK (CLI argv[1], 64..1024); 0 = default  */
  int32_t list_scalar;               /* 0/1: 4-byte list loads (default, fastest measured);
                                        2: 16-byte int4 list loads (experiment) */
  int64_t plane_stride;              /* LJ_SOA_D: doubles between planes (>= pn)           */
  int64_t row_begin, row_end;        /* update only rows [row_begin,row_end); 0,0 = all    */
  int64_t list_entries;              /* optional: entries allocated in `list` (number_of_pairs);
                                        0 = unknown.  Lets LJ_VARIANT_TILE_TMA round its bulk
                                        copies up to 16 bytes without reading past the array */
  int64_t ell_width;                 /* LJ_LIST_ELL_ROWS: entries per row of the padded table   */
  uint64_t mirror_token;             /* lj_list_mirror_token() of the mirror the caller vouches for: the
                                        cell-tile / cluster mirrors are used only when this matches the
                                        context's current mirror AND the arrays are the ones it was built
                                        from.  0 (a caller that knows nothing of mirrors, or one that has
                                        written other contents into the same arrays) = per-row kernels.  */
} lj_force_args;

LJ_API int lj_force_step(lj_ctx* ctx, const lj_force_args* args, void* stream);
/* The same step in two parts, for callers that overlap a ghost exchange with the force work
 * (z-slab decomposition): the list must have been built for a row range [row_begin,row_end) of
 * "owned" particles with LJ_LIST_TILES.  LJ_PART_INTERIOR runs the tiles none of whose stencil cell
 * layers held, at build time, a particle outside the row range -- they only read positions of owned
 * particles; LJ_PART_BOUNDARY runs the others and first refreshes the library's copy of the
 * positions outside the row range (the ghosts).  INTERIOR then BOUNDARY = one lj_force_step. */
enum { LJ_PART_ALL = 0, LJ_PART_INTERIOR = 1, LJ_PART_BOUNDARY = 2 };
LJ_API int lj_force_step_part(lj_ctx* ctx, const lj_force_args* args, int32_t part, void* stream);
/* `loop` back-to-back steps, the body of measure() (cuda/force_cuda.cu:333-335).  With
 * use_graph != 0 the steps are captured once into a CUDA graph and replayed. */
LJ_API int lj_force_loop(lj_ctx* ctx, const lj_force_args* args, int loop, int use_graph, void* stream);

/* ---------------------------------------------------------------- list build ---------- */
/* Replaces makepair() + register_pair() (cuda/force_cuda.cu:102-163) with an O(N) on-GPU
 * build: cell binning (counting sort) -> 27-cell stencil search -> prefix scan -> fill.
 * Semantics of the reference: open boundaries, listed iff r2 < search_len^2 (strict) with
 * r2 = fma(dz,dz,fma(dy,dy,dx*dx)) in FP64; full: every ordered pair i != j; half: i < j.
 * Output in the caller's numbering: number_of_partners[i], pointer = exclusive scan (pn
 * entries, no sentinel), sorted_list rows in a deterministic but unspecified order (or
 * ascending j with LJ_LIST_SORT_ROWS, which is what makepair() produces). */
enum {
  LJ_LIST_SORT_ROWS = 1,
  /* also build the library-owned CLUSTER PAIR LIST (union of 4 consecutive rows with member
   * masks) that LJ_VARIANT_CLUSTER / AUTO use: q[j] is gathered once per cluster entry and serves
   * up to four i-particles from registers.  Full lists only; one small host read-back per build.
   * The mirror is tied to the three output arrays: rebuilding into them, lj_shuffle_rows on them
   * or lj_list_invalidate() drops it. */
  LJ_LIST_CLUSTERS = 2,
  /* use the one-search-per-particle kernel instead of the default cluster-organised search
   * (identical output; kept for A/B measurements) */
  LJ_LIST_PER_PARTICLE_SEARCH = 4,
  /* also build the library-owned CELL-TILE MIRROR of the list: the same rows in cell order with
   * 16-bit indices into the shared-memory region of their tile (LJ_VARIANT_CELLTILE / AUTO).
   * Full lists, FP64 layouts; two small host read-backs per build.  Dropped like the cluster list;
   * silently absent when a tile's neighbourhood would not fit in shared memory (very dense
   * systems), in which case AUTO stays on the per-row kernels. */
  LJ_LIST_TILES = 8,
  /* With LJ_LIST_TILES: size the tiles of the mirror for the mixed-precision force kernel
   * (LJ_PREC_MIXED: 16-byte position records leave shared memory for ~72-row tiles, measured faster
   * than the ~56-row tiles the FP64 kernel prefers).  Either kernel runs on either mirror. */
  LJ_LIST_TILES_WIDE = 16
};

typedef struct lj_list_args {
  const void* q;                /* device, layout below (FP64 layouts only)               */
  int64_t pn;
  int32_t layout;
  int32_t half;                 /* 0 full list, 1 half list (EN_ACTION_REACTION build)    */
  int64_t plane_stride;
  double search_len;            /* SEARCH_LENGTH (3.3)                                    */
  int32_t* number_of_partners;  /* out, device int32[pn]                                  */
  void* pointer;                /* out, device int32[pn] or int64[pn]                     */
  int32_t* sorted_list;         /* out, device int32[capacity]                            */
  int64_t capacity;             /* entries available in sorted_list                       */
  int32_t pointer64;
  int32_t flags;
  int64_t row_begin, row_end;   /* build rows only for i in [row_begin,row_end); 0,0=all;
                                   all pn particles are neighbour candidates (ghosts)     */
} lj_list_args;

/* number_of_pairs_out (host, may be NULL): total entries; when non-NULL the call
 * synchronises the stream and returns LJ_ERR_CAPACITY (with the needed total stored) or
 * LJ_ERR_OVERFLOW32 instead of writing out of bounds (the reference silently overruns,
 * cuda/force_cuda.cu:102-120).  With NULL the call stays asynchronous; query later with
 * lj_list_result(). */
LJ_API int lj_build_list(lj_ctx* ctx, const lj_list_args* args, int64_t* number_of_pairs_out,
                  void* stream);
/* Build the cell-tile mirror for a CSR list the CALLER supplies -- built on the host, loaded from a
 * pair cache, shuffled (the reference's flow: cuda/force_cuda.cu:203-227, 255-263, 392-397 always hands
 * the kernel such a list).  Bins q, lays out the tiles and translates every entry j of row i into the
 * 16-bit index of j inside the shared-memory region of i's tile, in the row's own order (so the cell-tile
 * kernel sums in the same order as the per-row kernel with 8 lanes per row).  Rows with an entry outside
 * the region of their tile (a list built with a longer search length than `search_len`, or positions
 * that have moved further than the skin since) are left out of the mirror and served by the per-row
 * kernel on the caller's arrays right after the cell-tile kernel; *rows_outside_out tells how many.
 * Full lists, FP64 layouts.  Synchronises.  flags: LJ_LIST_TILES_WIDE or 0. */
LJ_API int lj_list_mirror(lj_ctx* ctx, const void* q, int64_t pn, int32_t layout, int64_t plane_stride,
                          double search_len, const int32_t* number_of_partners, const void* pointer,
                          int32_t pointer64, const int32_t* sorted_list, int64_t list_entries, int32_t flags,
                          int64_t* rows_outside_out, void* stream);
/* generation token of the mirror this context currently holds (0 = none): changes with every
 * lj_build_list / lj_list_mirror that produces one; pass it in lj_force_args.mirror_token */
LJ_API uint64_t lj_list_mirror_token(lj_ctx* ctx);
/* drop the cluster mirror (call after modifying the list arrays yourself) */
LJ_API int lj_list_invalidate(lj_ctx* ctx);
/* status + totals of the most recent lj_build_list on this context (synchronises `stream`) */
LJ_API int lj_list_result(lj_ctx* ctx, int64_t* number_of_pairs_out, int32_t* max_partners_out,
                   void* stream);

/* ------------------------------------------------ six-array SoA (OpenACC SoA program) - */
/* The SoA-on-GPU interface of openacc/force_oacc_soa.cpp: six separately allocated device
 * arrays qx,qy,qz,px,py,pz (:17-22) instead of one block with a plane stride.  `fa` / `la`
 * carry everything else (list arrays, dt, CL2, variant, ...); their q, p, layout and
 * plane_stride fields are ignored.  force_reactless (:203-232) = CSR list, force_reactless_memopt
 * (:234-263) = ELL list.  When the three q (and p) arrays are equally spaced in one allocation
 * the kernels read them in place (LJ_SOA_D); otherwise q is gathered into a library-owned SoA
 * block once per call, p is gathered before and scattered back after the `loop` steps, so the
 * copies amortise over the loop exactly like the reference's acc update device / update host
 * around its LOOP (:279-292). */
LJ_API int lj_force_loop_soa6(lj_ctx* ctx, const double* qx, const double* qy, const double* qz,
                              double* px, double* py, double* pz, const lj_force_args* fa, int loop,
                              void* stream);
LJ_API int lj_build_list_soa6(lj_ctx* ctx, const double* qx, const double* qy, const double* qz,
                              const lj_list_args* la, int64_t* number_of_pairs_out, void* stream);

/* Replaces make_transposed_pairlist() (cuda/force_cuda.cu:229-240): CSR -> column-major
 * ELL with stride pn, zero padding up to max_partners rows.  capacity_entries must be
 * >= max_partners*pn (LJ_ERR_CAPACITY otherwise; max_partners_out tells how many). */
LJ_API int lj_build_ell(lj_ctx* ctx, const int32_t* sorted_list, const int32_t* number_of_partners,
                 const void* pointer, int32_t pointer64, int64_t pn, int32_t* transposed_list,
                 int64_t capacity_entries, int32_t* max_partners_out, void* stream);

/* Replaces make_sorted_list2d() (cuda/force_cuda.cu:242-253) with a CORRECT padded table: CSR ->
 * row-major sorted_list2d[i*width + k], zero padded.  width must be >= the longest row
 * (LJ_ERR_CAPACITY otherwise -- the reference's NUM_NEIGH = 60 is smaller than its own max_partners
 * of 78 and its rows overlap silently; max_partners_out tells the width that is needed) and
 * capacity_entries >= width*pn.  Consumed by lj_force_step with LJ_LIST_ELL_ROWS + ell_width. */
LJ_API int lj_build_ell_rows(lj_ctx* ctx, const int32_t* sorted_list, const int32_t* number_of_partners,
                      const void* pointer, int32_t pointer64, int64_t pn, int32_t width,
                      int32_t* sorted_list2d, int64_t capacity_entries, int32_t* max_partners_out,
                      void* stream);

/* Replaces random_shfl() (cuda/force_cuda.cu:255-263) in spirit: a deterministic per-row
 * permutation on the device (NOT the same permutation as std::shuffle), to prove kernels do
 * not depend on row order. */
LJ_API int lj_shuffle_rows(lj_ctx* ctx, int32_t* sorted_list, const int32_t* number_of_partners,
                    const void* pointer, int32_t pointer64, int64_t pn, uint32_t seed,
                    void* stream);

/* Replaces check_loadedpair() (cuda/force_cuda.cu:183-201) on the device: 0 <= np < pn,
 * 0 <= pointer[i] <= number_of_pairs, 0 <= sorted_list[k] < pn.  Synchronises. */
LJ_API int lj_validate_list(lj_ctx* ctx, const int32_t* sorted_list, const int32_t* number_of_partners,
                     const void* pointer, int32_t pointer64, int64_t pn, int64_t number_of_pairs,
                     void* stream);

/* ---------------------------------------------------------------- host helpers -------- */
/* Replaces init() + add_particle() (cuda/force_cuda.cu:47-94): jittered FCC lattice, one
 * std::mt19937(2) stream, U(0,0.1) per coordinate, iz->iy->ix->basis order.  HOST function:
 * writes packed xyz doubles.  Returns the particle count, or -(needed) if cap is too small. */
LJ_API int64_t lj_init_fcc(double density, double L, double* q_xyz_host, int64_t cap_particles,
                    int32_t* cells_per_side_out);
/* Pair-list cache files of the reference, for exchanging lists with its binaries (host only).
 * Text `.cache_pair_{all,half}.dat`: makepaircache()/loadpair() of cuda/force_cuda.cu:165-227
 * (the reader applies check_loadedpair()'s range checks and the header pn check, :183-213).
 * Arrays may be NULL to query pn/npairs only; pn_expected < 0 accepts any particle count. */
LJ_API int lj_paircache_write_text(const char* path, int64_t pn, int64_t npairs,
                                   const int32_t* number_of_partners, const int32_t* pointer,
                                   const int32_t* sorted_list);
LJ_API int lj_paircache_read_text(const char* path, int64_t pn_expected, int64_t* pn_out,
                                  int64_t* npairs_out, int32_t* number_of_partners, int32_t* pointer,
                                  int64_t cap_particles, int32_t* sorted_list, int64_t cap_pairs);
/* Binary `pair.dat` of cpu_ref (savepair()/loadpair(), cpu_ref/force_soa.cpp:360-377):
 * int npairs; int number_of_partners[N]; int i_particles[MAX_PAIRS]; int j_particles[MAX_PAIRS];
 * n_static / max_pairs_static are the reference's compile-time N = 400000 and MAX_PAIRS = 30*N. */
LJ_API int lj_pairdat_write(const char* path, int64_t n_static, int64_t max_pairs_static, int64_t pn,
                            int64_t npairs, const int32_t* number_of_partners,
                            const int32_t* i_particles, const int32_t* j_particles);
LJ_API int lj_pairdat_read(const char* path, int64_t n_static, int64_t max_pairs_static, int64_t pn,
                           int64_t* npairs_out, int32_t* number_of_partners, int32_t* i_particles,
                           int32_t* j_particles, int64_t cap_pairs);
/* same lattice generated on the device for sizes where a host loop is too slow is NOT
 * offered: the generator is sequential by definition (one RNG stream). */

/* The reference's whole measure() (cuda/force_cuda.cu:319-342) as one call on HOST arrays:
 * upload q and p, build the neighbour list on the GPU (or upload the caller's), run `loop`
 * force steps rebuilding the list every `rebuild_every` steps (0 = never), download p.
 * This is the call bench.py times for its end-to-end number. */
typedef struct lj_measure_args {
  const void* q_host;
  void* p_host;                 /* in/out                                                 */
  int64_t pn;
  int32_t layout;
  int32_t half;                 /* 1: Newton-3 path on a half list                        */
  int64_t plane_stride;
  double dt, cl2, search_len;
  int32_t loop;                 /* LOOP (100)                                             */
  int32_t rebuild_every;        /* list rebuild cadence in steps; 0 = build once          */
  int32_t variant, group, precision, threads_per_block;
  int32_t use_graph;
  int32_t list_flags;
  /* optional caller-provided CSR list on the host (the reference uploads its host-built
   * list in copy_to_gpu, cuda/force_cuda.cu:302-312); NULL = build on the GPU */
  const int32_t* list_host;
  const int32_t* number_of_partners_host;
  const int32_t* pointer_host;
  int64_t number_of_pairs_in;
  /* outputs */
  int64_t number_of_pairs;      /* entries of the list the kernels consumed               */
  int32_t max_partners;
  int32_t list_builds;          /* how many GPU list builds ran                           */
  double seconds_total;         /* wall clock incl. H<->D (first stderr line of measure()) */
  double seconds_kernel;        /* "without Host<->Device" (second stderr line)           */
  int64_t h2d_bytes, d2h_bytes;
} lj_measure_args;

LJ_API int lj_measure(lj_ctx* ctx, lj_measure_args* args);

/* ---------------------------------------------------------------- MD step helpers ----- */
/* The caller the hot path is meant for (NOT in the reference, whose q is static): the
 * force call is the kick p += F dt of a symplectic Euler step, these add the drift, the
 * skin-based rebuild trigger and energies for conservation checks. */
LJ_API int lj_drift(lj_ctx* ctx, void* q, const void* p, int64_t pn, int32_t layout,
                    int64_t plane_stride, double dt, void* stream);           /* q += p dt */
/* max_i |q_i - q_ref,i|^2 (synchronises); rebuild the list when it exceeds ((search-cutoff)/2)^2 */
LJ_API int lj_max_displacement2(lj_ctx* ctx, const void* q, const void* q_ref, int64_t pn,
                                int32_t layout, int64_t plane_stride, double* out_host, void* stream);
/* kinetic = sum p^2/2, potential = sum over listed pairs with r2 <= cl2 of 4(r^-12 - r^-6)
 * (halved for a full list; pass variant = LJ_VARIANT_NEWTON3 for a half list).  Synchronises. */
LJ_API int lj_energy(lj_ctx* ctx, const lj_force_args* args, double* kinetic_out,
                     double* potential_out, void* stream);

/* ---------------------------------------------------------------- multi-GPU helpers --- */
/* z-slab decomposition support (no reference counterpart: the reference is single-GPU).
 * Peer access by CUDA IPC: export a handle for a device allocation, open a peer's. */
/* IPC-shareable device memory (plain cudaMalloc: the legacy IPC handles cannot describe
 * stream-ordered pool allocations).  Export/open work on pointers returned by lj_ipc_alloc. */
LJ_API int lj_ipc_alloc(lj_ctx* ctx, size_t bytes, void** out);
LJ_API int lj_ipc_free(lj_ctx* ctx, void* ptr);
LJ_API int lj_ipc_export(lj_ctx* ctx, void* dev_ptr, uint8_t handle_out[64]);
LJ_API int lj_ipc_open(lj_ctx* ctx, const uint8_t handle[64], void** peer_ptr_out);
LJ_API int lj_ipc_close(lj_ctx* ctx, void* peer_ptr);
/* Halo pull: copy `bytes` (multiple of 16) from a peer-mapped pointer into local memory
 * with a grid of 16-byte P2P loads over NVLink. */
LJ_API int lj_halo_pull(lj_ctx* ctx, void* local_dst, const void* peer_src, size_t bytes, void* stream);

/* ---------------------------------------------------------------- z-slab decomposition, one process --- */
/* The decomposed path for a C / C++ caller: ONE process drives `ngpus` devices, one lj_ctx each,
 * peer access between neighbours, no MPI / NCCL / torch.  The caller supplies the particles in an order in
 * which slab g is the contiguous index range [slab_begin[g], slab_begin[g+1]) and the ghosts a slab needs
 * are the last `halo_rows` particles of the slab below and the first `halo_rows` of the slab above -- what the
 * reference's generator produces (lattice layers z-outermost, cuda/force_cuda.cu:68-77); lj_decomp_plan_fcc
 * computes both for that lattice.  Each slab keeps [owned | ghosts below | ghosts above] as double4 on its
 * device, builds its list for the owned rows (LJ_LIST_TILES) and runs the cell-tile kernel in INTERIOR /
 * BOUNDARY parts around a device-ordered ghost pull (lj_flag_set, lj_halo_pull_sync, lj_flag_wait). */
typedef struct lj_decomp lj_decomp;
typedef struct lj_decomp_args {
  int32_t ngpus;
  const int32_t* devices;       /* CUDA ordinals, NULL = 0 .. ngpus-1                                   */
  const double* q_xyz_host;     /* [pn][3] positions, host                                              */
  int64_t pn;
  const int64_t* slab_begin;    /* [ngpus + 1], ascending, slab_begin[0] = 0, slab_begin[ngpus] = pn     */
  int64_t halo_rows;            /* particles exchanged per slab face                                     */
  double search_len, cutoff, dt; /* 0 = the reference's 3.3, 3.0, 0.001                                  */
  int32_t precision;            /* lj_precision                                                          */
  int32_t list_flags;           /* 0 = LJ_LIST_TILES (+ _WIDE for LJ_PREC_MIXED)                          */
} lj_decomp_args;
/* slab_begin[ngpus + 1] and halo_rows for the FCC lattice of lj_init_fcc(density, L): whole lattice layers
 * per slab; LJ_ERR_BAD_ARG when a slab would be thinner than the halo (its ghosts would live two slabs away) */
LJ_API int lj_decomp_plan_fcc(double density, double L, int32_t ngpus, double search_len, int64_t* slab_begin,
                              int64_t* halo_rows);
/* *out is set even on failure so that lj_decomp_last_error() can say why; destroy it either way */
LJ_API int lj_decomp_create(lj_decomp** out, const lj_decomp_args* args);
/* `nsteps` force steps on STATIC positions (the reference benchmark): the list is rebuilt every
 * `rebuild_every` steps (0 = never); overlap != 0: interior tiles run while the ghost pull is in flight.
 * Asynchronous: follow with lj_decomp_sync / lj_decomp_gather. */
LJ_API int lj_decomp_step(lj_decomp* d, int32_t nsteps, int32_t rebuild_every, int32_t overlap);
/* `nsteps` of kick (force step) + drift (q += p dt, owned particles), list rebuilt every `rebuild_every` steps */
LJ_API int lj_decomp_md(lj_decomp* d, int32_t nsteps, int32_t rebuild_every, int32_t overlap);
LJ_API int lj_decomp_rebuild(lj_decomp* d);
LJ_API int lj_decomp_sync(lj_decomp* d);
/* owned momenta (and positions, may be NULL) of all slabs in the caller's particle order, [pn][3] on the host */
LJ_API int lj_decomp_gather(lj_decomp* d, double* p_xyz_host, double* q_xyz_host);
LJ_API int64_t lj_decomp_pairs(lj_decomp* d);          /* list entries over all slabs */
LJ_API int64_t lj_decomp_launch_count(lj_decomp* d);   /* kernels launched by all its contexts */
LJ_API const char* lj_decomp_last_error(lj_decomp* d);
LJ_API int lj_decomp_destroy(lj_decomp* d);

/* Ordering between GPUs without host round trips: 32-bit step counters in peer-visible memory
 * (lj_ipc_alloc'ed, or any device memory a peer can address).  lj_flag_set stores `value` once all
 * earlier work of `stream` is done (system-scope release); lj_flag_wait holds `stream` until the flag
 * is >= at_least.  lj_halo_pull_sync copies one or two segments in one launch; each segment first
 * waits for *wait_flag >= wait_value (NULL: no wait; typically the owner's "q of step k is final")
 * and finally stores done_value into *done_flag (NULL: none; typically a counter in the OWNER's
 * memory, which the owner waits for before it overwrites q).  A wait only ever depends on an
 * earlier step of the other rank, so two ranks cannot wait for each other. */
typedef struct lj_halo_seg {
  void* local_dst;
  const void* peer_src;
  size_t bytes;                 /* multiple of 16 */
  const int32_t* wait_flag;
  int32_t wait_value;
  int32_t done_value;
  int32_t* done_flag;
} lj_halo_seg;
LJ_API int lj_flag_set(lj_ctx* ctx, int32_t* flag, int32_t value, void* stream);
LJ_API int lj_flag_wait(lj_ctx* ctx, const int32_t* flag, int32_t at_least, void* stream);
LJ_API int lj_halo_pull_sync(lj_ctx* ctx, const lj_halo_seg* segs, int32_t nsegs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LJ_B200_H */
