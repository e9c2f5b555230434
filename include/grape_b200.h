/* grape_b200.h — C ABI of the B200-native PIE graph engine.
 *
 * The reference (alibaba/libgrape-lite @ e7c4465) exposes NO C ABI: its GPU
 * path is header-only C++ compiled into each app.  This header is the thin
 * boundary the reference's C++ host would bind instead of grape/cuda/ **:
 * every entry point cites the reference interface it replaces
 * (paths relative to the reference root).  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns gl_status (0 = ok, <0 = error);
 *     gl_last_error() returns a thread-local message.  The C++ shim turns a
 *     non-zero status into LOG(FATAL), preserving the reference's
 *     abort-on-error behaviour (grape/cuda/utils/cuda_utils.h:60-108).
 *   - one host thread per device (mirrors one MPI rank per GPU,
 *     examples/analytical_apps/run_cuda_app.h:207-214); handles are not
 *     thread-safe.  The library uses the CUDA *current device* of the caller.
 *   - host pointers passed in descriptors are BORROWED for the call.
 *   - there is no CPU fallback: every call fails with GL_ERR_CUDA when no
 *     CUDA device is usable.
 */
#ifndef GRAPE_B200_H_
#define GRAPE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GL_ABI_VERSION 2

typedef enum {
  GL_OK = 0,
  GL_ERR_CUDA = -1,   /* CUDA runtime error / no device            */
  GL_ERR_ARG = -2,    /* invalid argument                          */
  GL_ERR_NOMEM = -3,  /* host or device allocation failed          */
  GL_ERR_STATE = -4,  /* call not valid in the handle's state      */
  GL_ERR_COMM = -5    /* peer / collective failure                 */
} gl_status;

const char* gl_last_error(void);
int gl_abi_version(void);
/* Device facts used for grid sizing: sm_count, cc major*10+minor, L2 bytes. */
int gl_device_info(int* sm_count, int* cc, size_t* l2_bytes, size_t* hbm_bytes);

/* ------------------------------------------------------------------ *
 * Load balancing of ForEachEdge (grape/cuda/parallel/parallel_engine.h:51,
 * ParseLoadBalancing :53-70; flag --lb, examples/analytical_apps/flags.cc).
 * ------------------------------------------------------------------ */
typedef enum {
  GL_LB_NONE = 0,   /* thread per vertex            (LBNONE   :621-646) */
  GL_LB_CM = 1,     /* CTA-cooperative mapping      (LBCM     :716-771) */
  GL_LB_WM = 2,     /* warp-cooperative mapping     (LBWARP   :773-845) */
  GL_LB_CTA = 3,    /* 3-tier CTA/warp/thread       (LBCTA    :847-879) */
  GL_LB_STRICT = 4, /* exact edge-balanced          (LBSTRICT :881-979) */
  GL_LB_CMOLD = 5   /* alias of CM (LBCMOld :648-714)                   */
} gl_lb;

/* grape::LoadStrategy (grape/types.h) */
typedef enum { GL_LOAD_ONLY_OUT = 0, GL_LOAD_BOTH_OUT_IN = 1 } gl_load_strategy;

/* ------------------------------------------------------------------ *
 * Fragment: replaces grape::cuda::HostFragment / dev::DeviceFragment
 * (grape/cuda/fragment/host_fragment.h:100-120,322-438,
 *  grape/cuda/fragment/device_fragment.h:36-450).
 * Storage is SoA: 64-bit row pointers, 32-bit neighbour lids, separate
 * weight array; rows sorted by neighbour lid, inner neighbours first
 * (grape/graph/immutable_csr.h:104-131).
 * ------------------------------------------------------------------ */
typedef struct gl_frag gl_frag_t;

/* One CSR as the host's ImmutableEdgecutFragment holds it
 * (grape/fragment/csr_edgecut_fragment_base.h: oe_/ie_ + offsets). */
typedef struct {
  const uint64_t* row_ptr; /* rows+1 entries                                 */
  const uint32_t* col;     /* neighbour lid                                  */
  const void* edata;       /* NULL | float[] | double[] (see edata_bytes)    */
  uint64_t rows;           /* number of rows described                       */
} gl_csr_desc;

typedef struct {
  uint32_t fid, fnum;
  int directed;
  int load_strategy;   /* gl_load_strategy                                   */
  uint32_t ivnum;      /* inner vertices: lids [0, ivnum)                     */
  uint32_t ovnum;      /* outer vertices: lids [ivnum, ivnum+ovnum)           */
  uint64_t total_vnum; /* GetTotalVerticesNum()                               */
  int edata_bytes;     /* 0 (EmptyType), 4 (float) or 8 (double)              */
  gl_csr_desc oe;      /* rows = ivnum (inner rows; outer rows are empty)     */
  gl_csr_desc ie;      /* directed: rows = ivnum; undirected: ignored         */
  gl_csr_desc ov_ie;   /* optional reverse adjacency of OUTER vertices, rows =
                          ovnum, cols = inner lids (what the reference keeps
                          in ie_ rows ivnum..tvnum for kOnlyOut undirected,
                          csr_edgecut_fragment_base.h:475-492); rows==0 -> the
                          library derives it from oe                          */
  const uint32_t* ovgid;     /* gid of each outer vertex, ascending (IdParser
                                format, grape/fragment/id_parser.h:28-55)     */
  const int64_t* inner_oids; /* oid of each inner vertex, ascending; NULL ->
                                oid = oid_base + lid                          */
  int64_t oid_base;
} gl_frag_desc;

/* replaces HostFragment::__allocate_device_fragment__ (host_fragment.h:322-438).
 * Preconditions the library relies on (the reference's loaders satisfy them):
 *  - rows sorted by neighbour lid, inner neighbours first; ovgid ascending;
 *  - a DIRECTED fragment passes its in-edges in d->ie (without them BFS only pushes, PageRank pull
 *    falls back to push and WCC / wcc_opt are refused);
 *  - lids were assigned in ascending oid order inside each fragment and fids in ascending oid
 *    blocks IF the smallest-label tie-breaks of CDLP / WCC must equal the reference's smallest-OID
 *    rule (the apps tie-break on the smallest GID); with another vertex map attach a gl_vm_t
 *    (gl_app_set_vertex_map) to get oids back and expect ties to follow gid order. */
int gl_frag_create(gl_frag_t** out, const gl_frag_desc* d);
/* HostFragment::PrepareToRunApp (host_fragment.h:122-215): split positions, outer ranges, the
 * reverse adjacency of the outer copies and the destination-fragment information are part of the
 * device layout from gl_frag_create on, so this only validates the request: need_mirror_info needs a
 * communicator with a mirror area at sync time, need_build_device_vm is served by gl_vm_create. */
int gl_frag_prepare(gl_frag_t*, int message_strategy, int need_split_edges, int need_mirror_info);

/* Edge-list front end: replaces LoadGraph -> EVFragmentLoader ->
 * ImmutableEdgecutFragment::Init -> buildCSR -> upload
 * (grape/fragment/loader.h:46-53, immutable_edgecut_fragment.h:215-350,
 * csr_edgecut_fragment_base.h:417-734) with a device-side build (radix sort).
 * Partitioner: contiguous blocks of ceil(n/fnum) ascending oids
 * (grape/vertex_map/partitioner.h:115-126 SegmentedPartitioner); lid = rank
 * of the oid inside the block, so gid order == oid order.
 * Vertices are oids 0..n-1 (oids==NULL) or the given ascending oid list.
 * Every rank passes the edges it holds; edges with no inner endpoint are
 * ignored, edges naming unknown oids are dropped. */
typedef struct {
  uint64_t n_vertices;
  const int64_t* oids;  /* NULL or ascending list of n_vertices oids          */
  uint64_t n_edges;
  const int64_t* src;   /* host pointers                                      */
  const int64_t* dst;
  const void* edata;    /* NULL | float[] | double[]                          */
  int edata_bytes;      /* 0 | 4 | 8 ; device copy keeps this width           */
  int directed;
  int load_strategy;
  uint32_t fid, fnum;
} gl_edges_desc;
int gl_frag_build_from_edges(gl_frag_t** out, const gl_edges_desc* d);

/* Synthetic Graph500 R-MAT fragment built entirely on the device
 * (A,B,C,D = .57,.19,.19,.05; ids scrambled by a fixed
 * bijection; self loops and duplicates kept; undirected).  weight_mode:
 * 0 none, 1 integer-valued 1..255 stored as f32, 2 real (0,1] stored as f32. */
typedef struct {
  int scale;
  int edgefactor;
  uint64_t seed;
  int weight_mode;
  uint32_t fid, fnum;
} gl_rmat_desc;
int gl_frag_build_rmat(gl_frag_t** out, const gl_rmat_desc* d);
/* Host copy of the same generator (inputs for the CPU oracle / reference);
 * w may be NULL. Generates edges [first, first+count). */
int gl_rmat_edges_host(const gl_rmat_desc* d, uint64_t first, uint64_t count,
                       int64_t* src, int64_t* dst, float* w);

typedef struct {
  uint32_t fid, fnum, ivnum, ovnum;
  uint64_t total_vnum;
  uint64_t oe_num, ie_num; /* CSR entries held (GetOutgoingEdgeNum etc.)      */
  int directed, load_strategy, edata_bytes;
  int fid_offset;          /* IdParser::fid_offset_                           */
  uint64_t device_bytes;   /* HBM held by the fragment                        */
  uint32_t max_degree;
} gl_frag_info;
int gl_frag_get_info(const gl_frag_t*, gl_frag_info* out);

/* Device view (POD of device pointers): replaces HostFragment::DeviceObject()
 * (host_fragment.h:277-320).  Invalidated by destroy; while the topology is
 * offloaded oe_col / oe_w (and their ie aliases) are NULL, the id fields stay valid. */
typedef struct {
  uint32_t fid, fnum, ivnum, ovnum;
  uint64_t total_vnum;
  int directed, edata_bytes, fid_offset;
  uint32_t id_mask;
  const uint64_t* oe_rp;    /* [ivnum+1] */
  const uint32_t* oe_col;
  const void* oe_w;
  const uint64_t* oe_split; /* [ivnum] first outer-neighbour position of the row
                               (oespliters_, immutable_edgecut_fragment.h:744-772) */
  const uint64_t* ie_rp;    /* directed: [ivnum+1]; undirected: == oe_rp      */
  const uint32_t* ie_col;
  const void* ie_w;
  const uint64_t* ie_split;
  const uint64_t* ovie_rp;  /* [ovnum+1] reverse adjacency of outer vertices   */
  const uint32_t* ovie_col; /* inner lids                                     */
  const uint32_t* ovgid;    /* [ovnum]                                        */
  const uint32_t* outer_range; /* [fnum+1] outer lids owned by each fid
                                  (outer_vertices_of_frag_)                   */
  const int64_t* inner_oids;   /* may be NULL (oid = oid_base + lid)          */
  int64_t oid_base;
  uint64_t oe_num, ie_num;     /* CSR entries behind oe_col / ie_col          */
} gl_frag_view;
int gl_frag_view_get(const gl_frag_t*, gl_frag_view* out);

/* D2H copies of the stored CSR (layout parity tests; Serialize analogue,
 * grape/graph/immutable_csr.h:307-363).  which: 0 = oe, 1 = ie, 2 = ov_ie.
 * Any of the output pointers may be NULL. w is written with edata_bytes width. */
int gl_frag_copy_csr(const gl_frag_t*, int which, uint64_t* row_ptr,
                     uint32_t* col, void* w);
int gl_frag_copy_ovgid(const gl_frag_t*, uint32_t* ovgid);
/* GetInnerVertex(oid) -> lid (host; host_fragment / fragment_base API).
 * Returns GL_OK and *lid, or GL_ERR_ARG when the oid is not an inner vertex. */
int gl_frag_oid2lid(const gl_frag_t*, int64_t oid, uint32_t* lid);
/* max-out-degree inner vertex (ties -> smallest lid) and its degree */
int gl_frag_max_degree_vertex(const gl_frag_t*, uint32_t* lid, uint64_t* degree);
/* Binary cache of a fragment: ImmutableEdgecutFragment::Serialize / Deserialize
 * (grape/fragment/immutable_edgecut_fragment.h:508-584, immutable_csr.h:307-363;
 * the reference writes <prefix>/frag_<fid>.s).  Header + raw SoA arrays; a load
 * is fread + H2D, nothing is re-sorted. */
int gl_frag_save(const gl_frag_t*, const char* path);
int gl_frag_load(gl_frag_t** out, const char* path);
/* OffloadTopology / ReloadTopology (host_fragment.h:440-468) */
int gl_frag_offload(gl_frag_t*);
int gl_frag_reload(gl_frag_t*);
void gl_frag_destroy(gl_frag_t*);

/* ------------------------------------------------------------------ *
 * Device vertex map (oid <-> gid): replaces grape::cuda::DeviceVertexMap
 * (grape/cuda/vertex_map/device_vertex_map.h:94-172; built when an app sets
 * need_build_device_vm, host_fragment.h:205-215).  One device array of oids
 * per group (fragment f at off[f]) + binary search instead of the reference's
 * chained hash maps.  oids[f] = host array of fragment f's inner oids in lid
 * order (VertexMap::GetOid(fid, lid)).
 * ------------------------------------------------------------------ */
typedef struct gl_vm gl_vm_t;
typedef struct {
  uint32_t fnum;
  int fid_offset;
  uint32_t id_mask;
  const int64_t* l2o;          /* device [off[fnum]]: oid of (f, lid) at off[f] + lid   */
  const uint64_t* off;         /* device [fnum+1]                                        */
  const int64_t* sorted_oid;   /* device, NULL when every slice of l2o is ascending      */
  const uint32_t* sorted_lid;  /* device, lid of sorted_oid[i]                           */
} gl_vm_view;
int gl_vm_create(gl_vm_t** out, uint32_t fnum, const uint64_t* ivnums, const int64_t* const* oids);
int gl_vm_view_get(const gl_vm_t*, gl_vm_view* out);          /* DeviceObject()          */
/* batch lookups on device arrays: dev::DeviceVertexMap::GetGid / GetOid (:36-80);
 * unknown oid -> gid 0xFFFFFFFF, invalid gid -> oid -1 */
int gl_vm_oid2gid(const gl_vm_t*, void* stream, const int64_t* d_oids, uint64_t n, uint32_t* d_gids);
int gl_vm_gid2oid(const gl_vm_t*, void* stream, const uint32_t* d_gids, uint64_t n, int64_t* d_oids);
void gl_vm_destroy(gl_vm_t*);

/* ------------------------------------------------------------------ *
 * Fragment group communicator: replaces GPUMessageManager's NCCL/MPI
 * bootstrap (grape/cuda/parallel/gpu_message_manager.h:160-194) and
 * cuda::Communicator (grape/cuda/communication/communicator.h:41-95).
 * The data plane is NVLink peer memory: each rank exports its landing
 * buffers as CUDA IPC handles (gl_comm_export), the launcher exchanges the
 * 64-byte handles (any out-of-band channel: torch.distributed, a file) and
 * every rank maps its peers (gl_comm_open).  Scalar collectives and the
 * per-superstep barrier go through a caller-supplied callback so the
 * library needs no MPI/NCCL link dependency.
 * ------------------------------------------------------------------ */
typedef struct gl_comm gl_comm_t;
/* sum/min/max all-reduce of n int64 / double values in place (host memory). */
typedef int (*gl_allreduce_fn)(void* user, void* inout, int n, int is_double,
                               int op /*0 sum,1 min,2 max*/);
typedef struct {
  uint32_t fid, fnum;
  gl_allreduce_fn allreduce; /* required when fnum > 1 */
  void* user;
  size_t landing_bytes;      /* per-peer landing buffer capacity (bytes)      */
  size_t mirror_bytes;       /* per-peer capacity of the dense mirror-sync area
                                (inner state -> outer copies, the reference's
                                BatchShuffleMessageManager::SyncInnerVertices,
                                grape/cuda/parallel/batch_shuffle_message_manager.h:142-220);
                                0 = 8 bytes per inner vertex of the largest fragment
                                is NOT assumed: pass it explicitly                */
} gl_comm_desc;
#define GL_IPC_HANDLE_BYTES 64
int gl_comm_create(gl_comm_t** out, const gl_comm_desc* d);
/* writes fnum_slots*GL_IPC_HANDLE_BYTES bytes: handle of this rank's landing
 * area (one allocation, sliced per source rank) + control block */
int gl_comm_export(gl_comm_t*, void* handles_out, size_t bytes);
/* all_handles: fnum consecutive exports, ordered by fid */
int gl_comm_open(gl_comm_t*, const void* all_handles, size_t bytes);
/* Unmaps the peers' landing areas (keeps the local one).  Teardown order for a
 * group: every rank closes its peers, the group meets in a barrier, then each
 * rank destroys its communicator -- nobody frees memory a peer still maps. */
int gl_comm_close_peers(gl_comm_t*);
void gl_comm_destroy(gl_comm_t*);
/* Diagnostic: average time (us) of one kernel that stores `bytes` from local
 * memory into the next peer's (all_peers = 0) or every peer's (1) mirror slot
 * with 16-byte (vec16 = 1) or 4-byte stores, followed by a fence.sys per CTA.
 * Sizes the multi-GPU design decisions in DESIGN.md section 5. */
int gl_comm_peer_write_us(gl_comm_t*, size_t bytes, int vec16, int all_peers, int reps, double* us_out);

/* ------------------------------------------------------------------ *
 * Message manager (Face 2): the halo exchange that the shimmed
 * grape::cuda::GPUMessageManager / dev::MessageManager sit on
 * (grape/cuda/parallel/gpu_message_manager.h:45-458,
 *  grape/cuda/serialization/in_archive.h:36-169, out_archive.h:32-178,
 *  grape/cuda/parallel/message_kernels.h:28-127).
 * A producer kernel appends bytes for fragment d at
 *   send_slot[d] + atomicAdd(&send_bytes[d], sizeof(item))
 * -- send_slot[d] is fragment d's landing slot for this rank, mapped over
 * NVLink / CUDA IPC: the store IS the transfer (no staging archive, no NCCL
 * send/recv, no host size exchange).  gl_mm_finish_round publishes the byte
 * counts, runs the device-side barrier + termination vote and flips the
 * double buffer; what was sent in round r is readable through recv_slot /
 * recv_bytes during round r+1 only (the reference's rule: messages are
 * visible in the next IncEval).  Termination = no fragment sent a byte and
 * none called force_continue (gpu_message_manager.h:413-430).
 * ------------------------------------------------------------------ */
typedef struct gl_mm gl_mm_t;
typedef struct {
  uint32_t fid, fnum;
  int fid_offset;            /* gid = fid << fid_offset | lid (IdParser)        */
  uint32_t id_mask;
  uint32_t capacity_bytes;   /* per landing slot                               */
  char* const* send_slot;    /* device table [fnum]: my slot at fragment d     */
  uint32_t* send_bytes;      /* device [fnum]: bytes appended this round       */
  const char* const* recv_slot;   /* device table [fnum]: what fragment s sent me last round */
  const uint32_t* recv_bytes;     /* device [fnum]                              */
} gl_mm_view;
/* GPUMessageManager::Init (:160-194) on an OPENED communicator; fnum == 1: comm may be NULL */
int gl_mm_create(gl_mm_t** out, gl_comm_t* comm);
/* InitBuffer (:196-204): checks the per-peer capacities against the communicator's landing slots */
int gl_mm_init_buffer(gl_mm_t*, size_t send_bytes_per_peer, size_t recv_bytes_per_peer);
int gl_mm_start(gl_mm_t*);                         /* Start (:219) + round counter reset */
int gl_mm_start_round(gl_mm_t*, void* stream);     /* StartARound (:225-231)             */
int gl_mm_finish_round(gl_mm_t*, void* stream);    /* FinishARound (:237-303) + syncLengths (:400-431); synchronises the stream */
int gl_mm_to_terminate(gl_mm_t*, int* out);        /* ToTerminate (:321)                 */
int gl_mm_force_continue(gl_mm_t*);                /* ForceContinue (:338)               */
int gl_mm_view_get(gl_mm_t*, gl_mm_view* out);     /* DeviceObject (:342-344): valid for the current round */
uint64_t gl_mm_bytes_sent(gl_mm_t*);               /* GetMsgSize, accumulated over the query */
void gl_mm_destroy(gl_mm_t*);

/* Fixed-function consumers of ParallelProcess (:362-393): the received bytes
 * are (uint32 gid, value) pairs -- thrust::pair<vid_t, MESSAGE_T> -- or bare
 * uint32 gids (GL_MSG_SET_BIT); lid = gid & id_mask. */
typedef enum {
  GL_MSG_SET_BIT = 0,   /* bare gid: out_bitmap |= bit(lid)                     */
  GL_MSG_MIN_U32 = 1,   /* if (v < atomicMin(state[lid], v)) out_bitmap |= bit   */
  GL_MSG_MIN_F32 = 2,
  GL_MSG_MIN_F64 = 3,
  GL_MSG_ADD_F32 = 4,   /* atomicAdd(state[lid], v)                              */
  GL_MSG_ADD_F64 = 5
} gl_msg_op_kind;
typedef struct {
  int kind;               /* gl_msg_op_kind                                     */
  void* state;            /* device array indexed by lid                        */
  uint32_t* out_bitmap;   /* may be NULL                                        */
} gl_msg_op;
int gl_mm_process(gl_mm_t*, void* stream, const gl_msg_op* op, uint64_t* items_host /* may be NULL */);

/* Fixed-function producer: the "ForEach over outer vertices +
 * SyncStateOnOuterVertex[WarpOpt]" idiom (e.g. cuda/sssp/sssp.h:295-304).  For
 * every set bit v of `remote` inside the fragment's outer range, appends
 * (gid(v)[, state[v]]) to the owner's landing slot (value_bytes: 0 = bare gid,
 * 4 or 8 = thrust::pair<vid_t, value> layout) and clears the bit when
 * clear_bits != 0. */
int gl_mm_send_outer(gl_mm_t*, void* stream, const gl_frag_t* frag, uint32_t* remote_bitmap,
                     const void* state, int value_bytes, int clear_bits);

/* Dense mirror sync: owner state -> the outer copies other fragments hold
 * (BatchShuffleMessageManager::SyncInnerVertices,
 * grape/cuda/parallel/batch_shuffle_message_manager.h:142-220; mirror lists =
 * edgecut_fragment_base.h:569-603).  The communicator must have been created
 * with mirror_bytes >= 8 * (largest inner vertex count of the group).
 * gl_mm_mirror_plan is collective and runs once per (mm, fragment); every
 * sync is collective: owners pack their mirrored values straight into the
 * holders' mirror slots (peer stores), a device-side barrier follows, holders
 * copy the slots into their outer range.  values: device array over ALL local
 * vertices (inner then outer), elem_bytes 4 or 8; bitmap: one bit per local
 * vertex (ghost bits |= owner bits). */
int gl_mm_mirror_plan(gl_mm_t*, void* stream, const gl_frag_t* frag);
int gl_mm_sync_values_to_ghosts(gl_mm_t*, void* stream, void* values, int elem_bytes);
int gl_mm_sync_bits_to_ghosts(gl_mm_t*, void* stream, uint32_t* bitmap);

/* cuda::Communicator::Sum/Min/Max on one host scalar
 * (grape/cuda/communication/communicator.h:41-84,160-172): device-side peer
 * all-reduce, bit-identical on every rank (fixed fid order).
 * dtype: 0 int64, 1 double; op: 0 sum, 1 min, 2 max.  Synchronises `stream`. */
int gl_allreduce(gl_mm_t*, void* stream, void* inout_host, int dtype, int op);

/* ------------------------------------------------------------------ *
 * PIE apps (PEval / IncEval / Output): replaces GPUWorker::{Init,Query}
 * (grape/cuda/worker/gpu_worker.h:44-107) running the six GPU apps of
 * examples/analytical_apps/cuda/{bfs,sssp,wcc,pagerank,cdlp,lcc}.
 * ------------------------------------------------------------------ */
typedef enum {
  GL_APP_BFS = 0,      /* cuda/bfs/bfs.h        result: int64 depth            */
  GL_APP_SSSP = 1,     /* cuda/sssp/sssp.h      result: double (f32/f64 math)  */
  GL_APP_WCC = 2,      /* cuda/wcc/wcc.h        result: int64 oid of min-gid   */
  GL_APP_PAGERANK = 3, /* cuda/pagerank/pagerank.h result: double              */
  GL_APP_CDLP = 4,     /* cuda/cdlp/cdlp.h      result: int64 label            */
  GL_APP_LCC = 5,      /* cuda/lcc/lcc_opt.h    result: double                 */
  GL_APP_WCC_OPT = 6   /* cuda/wcc/wcc_opt.h (union-find) result: as GL_APP_WCC  */
} gl_app_kind;

typedef struct {
  int lb;                /* gl_lb (AppConfig::lb, cuda/app_config.h:22-27)     */
  int64_t source_oid;    /* --bfs_source / --sssp_source                        */
  double pr_delta;       /* --pr_d                                              */
  int max_round;         /* --pr_mr / --cdlp_mr                                 */
  int sssp_f64;          /* 1: fp64 distances on device (real-weight 1e-6 run)  */
  double sssp_prio;      /* 0 -> reference heuristic 32*avg_w/avg_deg           */
  int direction_opt;     /* BFS: 0 push only, 1 push/pull (bfs.h:168-261)       */
  int pr_pull;           /* PageRank: 0 push (atomicAdd), 1 pull (deterministic)*/
  int fuse_supersteps;   /* BFS: 1 = the whole query is ONE cooperative kernel per
                            GPU (levels, direction switches and, on several
                            fragments, the NVLink collectives happen inside it);
                            0 = one superstep per host round                    */
  int reserved[8];       /* tuning / test hooks, 0 = default:
                            [0] 1: WCC dense rounds as pull sweeps
                            [1] 1: BFS without the hub-first shadow CSR; 2: with it also on a
                                   small graph (default: from 2^16 vertices on); 3: like 2, several
                                   fragments, without the delegated-hub lists (A/B)
                            [2] BFS pull->push threshold divisor (default 24)
                            [3] >=4: number of BFS level bitmaps (forces spills)
                            [4] 1: PageRank f32 pull without the shared-memory hub table (A/B)
                            [5] 1: PageRank pull gathers f32 contributions
                            [7] 1: fused multi-fragment BFS ships per-holder bit-compressed
                                   frontier slices (round-1 scheme) instead of replicating
                                   the global frontier bitmap
                            [6] 1: BFS result as an int64 device array + one D2H
                                   (default: u8 depths over PCIe, widened on the host) */
} gl_app_config;
void gl_app_config_default(gl_app_config*);

typedef struct gl_app gl_app_t;
/* GPUWorker::Init: PrepareToRunApp + context Init (gpu_worker.h:44-65) */
int gl_app_create(gl_app_t** out, int kind, gl_frag_t* frag, gl_comm_t* comm,
                  const gl_app_config* cfg);

#define GL_MAX_STEP_STATS 512
typedef struct {
  int supersteps;            /* rounds executed (PEval counts as round 0)      */
  double query_ms;           /* device time of the whole Query()               */
  uint64_t entries_scanned;  /* CSR entries read by edge-scan kernels          */
  uint64_t frontier_vertices;/* sum over supersteps                            */
  uint64_t touched_vertices; /* distinct-destination updates (state writes)    */
  uint64_t kernel_launches;  /* kernels launched inside Query()                */
  uint64_t msg_bytes_sent;   /* halo bytes written to peers                    */
  int n_steps;               /* valid entries in the per-step arrays           */
  float step_ms[GL_MAX_STEP_STATS];
  uint64_t step_entries[GL_MAX_STEP_STATS];
  uint32_t step_frontier[GL_MAX_STEP_STATS];
  uint8_t step_mode[GL_MAX_STEP_STATS]; /* 0 push, 1 pull, 2 dense            */
} gl_query_stats;

/* GPUWorker::Query: PEval then IncEval until every fragment is idle
 * (gpu_worker.h:69-107). Asynchronous work is finished when it returns.
 * Collective on a fragment group: every rank calls it for the same query.  With
 * an opened communicator the call first meets the other ranks in a device-side
 * barrier on the app's stream (the MPI_Barrier in front of the reference's timed
 * Query(), run_cuda_app.h:117); query_ms starts after it.
 * step_frontier[] of a several-fragment fused BFS is THIS fragment's share of
 * each level's frontier. */
int gl_app_query(gl_app_t*, gl_query_stats* stats /* may be NULL */);
/* Context::Output values for the inner vertices, in lid order, to HOST memory
 * (e.g. cuda/sssp/sssp.h:104-118).  elem: see gl_app_kind. */
int gl_app_result(gl_app_t*, void* host_out, size_t bytes);
/* WCC / CDLP keep labels as gids; with a vertex map attached, gl_app_result
 * returns them as oids for ANY partition (without one: only for this library's
 * contiguous-block builders, where gid -> oid is arithmetic or the fragment's
 * own oid list).  The map is borrowed: keep it alive while the app is used. */
int gl_app_set_vertex_map(gl_app_t*, const gl_vm_t* vm);
/* oid of each inner vertex (first column of Output) */
int gl_app_result_oids(gl_app_t*, int64_t* host_out, size_t count);
void gl_app_destroy(gl_app_t*);

/* ------------------------------------------------------------------ *
 * Engine primitives: fixed-function fast paths of ParallelEngine
 * (grape/cuda/parallel/parallel_engine.h:72-293,987-1013) and the frontier
 * containers of grape/cuda/utils/{bitset,vertex_set,queue}.h.
 * Pointers below are DEVICE pointers owned by the caller.
 * ------------------------------------------------------------------ */
typedef enum {
  GL_OP_BFS_LEVEL = 0,    /* if level[v]==INF: level[v]=depth; out.insert(v)  */
  GL_OP_MIN_RELAX_U32 = 1,/* atomicMin(state[v], state[u] (+w))               */
  GL_OP_MIN_RELAX_F32 = 2,
  GL_OP_ADD_SCATTER_F64 = 3,
  GL_OP_COUNT = 4         /* counts scanned entries only (bandwidth probe)    */
} gl_edge_op_kind;

typedef struct {
  int kind;               /* gl_edge_op_kind                                   */
  void* state;            /* per-vertex state array (u32 / f32 / f64)          */
  void* state2;           /* ADD_SCATTER: destination array                    */
  uint32_t* out_bitmap;   /* updated vertices (atomicOr), may be NULL          */
  uint32_t depth;         /* BFS_LEVEL: level to write                         */
  int use_weight;
} gl_edge_op;

/* ForEachOutgoingEdge(stream, frag, WorkSourceArray(queue,n), op, lb)
 * (parallel_engine.h:987-1013).  queue: device array of n lids. */
int gl_edge_scan_queue(const gl_frag_t*, void* stream, const uint32_t* queue,
                       uint32_t n, const gl_edge_op* op, int lb,
                       uint64_t* entries_scanned_host /* may be NULL */);
/* bitmap (n_bits) -> queue of set bit positions; returns count on the host.
 * Replaces the O(V) ForEach + AppendWarp idiom (cuda/sssp/sssp.h:223-232). */
int gl_compact_bitmap(void* stream, const uint32_t* bitmap, uint32_t n_bits,
                      uint32_t* queue_out, uint32_t* count_host);

/* Frontier containers (grape/cuda/utils/bitset.h:32-196, vertex_set.h:30-163):
 * a DenseVertexSet is a device bitmap of ceil(nbits/32) words; VertexArray is
 * a plain device array (gl_dev_alloc + gl_dev_h2d / gl_dev_d2h); Queue = the
 * (queue_out, count) pair gl_compact_bitmap fills. */
int gl_bitmap_create(uint32_t** out, uint64_t nbits);               /* zeroed */
int gl_bitmap_clear(void* stream, uint32_t* bitmap, uint64_t nbits);  /* Clear(stream) */
int gl_bitmap_count(void* stream, const uint32_t* bitmap, uint64_t nbits, uint64_t* count_host); /* Count(stream): syncs */
int gl_bitmap_destroy(uint32_t* bitmap);

/* Queue<T> (grape/cuda/utils/queue.h:30-140): a device array of vertex ids + its length on the device. */
typedef struct gl_queue gl_queue_t;
int gl_queue_create(gl_queue_t** out, uint32_t capacity);
int gl_queue_clear(gl_queue_t*, void* stream);                          /* Clear(stream)            */
int gl_queue_size(gl_queue_t*, void* stream, uint32_t* size_host);      /* size(stream): syncs      */
int gl_queue_data(gl_queue_t*, uint32_t** data_dev, uint32_t** count_dev);  /* DeviceObject()       */
/* bitmap -> queue without the host round trip of gl_compact_bitmap (count stays on the device) */
int gl_queue_fill_from_bitmap(gl_queue_t*, void* stream, const uint32_t* bitmap, uint32_t n_bits);
void gl_queue_destroy(gl_queue_t*);
/* VertexArray<T> (grape/cuda/utils/vertex_array.h:34-169): device array + H2D / D2H */
int gl_varray_create(void** out, uint64_t count, int elem_bytes, int fill_byte);
int gl_varray_h2d(void* varray, const void* host, uint64_t count, int elem_bytes);
int gl_varray_d2h(const void* varray, void* host, uint64_t count, int elem_bytes);
int gl_varray_destroy(void* varray);

/* device scratch helpers so a C/ctypes caller can run the primitives */
int gl_dev_alloc(void** out, size_t bytes);
int gl_dev_free(void* p);
int gl_dev_memset(void* p, int byte, size_t bytes);
int gl_dev_h2d(void* dst, const void* src, size_t bytes);
int gl_dev_d2h(void* dst, const void* src, size_t bytes);
int gl_dev_sync(void);
/* page-locked host memory for result/ingest buffers (cudaMallocHost) */
int gl_host_alloc_pinned(void** out, size_t bytes);
int gl_host_free_pinned(void* p);
/* kernels launched by this library on the calling thread so far */
uint64_t gl_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GRAPE_B200_H_ */
