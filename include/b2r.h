/*
 * b2r.h - C ABI of the B200-native per-frame rendering path (libb2r.so).
 *
 * The reference (kdedeurw/Custom-GPGPU-Rasterizer) has no plugin / FFI interface: its
 * renderer is a C++ class inside one executable.  This header is the boundary a
 * maintainer of the reference binds instead of its CUDA translation units: each entry
 * point names the reference interface it replaces (paths relative to
 * CustomCUDARasterizer/).  The C++ veneer in custom_gpgpu_rasterizer_b200/csrc/host/ keeps the
 * reference's own class surface (CUDARenderer, SceneManager, SceneGraph, Mesh, ObjParser,
 * Camera, CUDATextureManager ...) and is implemented on top of these calls only.
 *
 * Conventions
 *   - every call returns B2R_OK (0) or a negative B2R_E_* code; b2r_last_error() gives text.
 *   - one context = one CUDA device + one stream.  A context is thread-compatible
 *     (no internal locking; do not call into one context from two threads at once).
 *   - matrices are 16 floats in the reference's FMatrix4 memory layout, data[col][row]
 *     (Matrix4.h:17): element (row r, col c) is m[c*4 + r].
 *   - host arrays passed in stay owned by the caller; device copies are owned by the context.
 *   - there is no CPU fallback: without a CUDA device b2r_create fails.
 */
#ifndef B2R_H
#define B2R_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2r_context b2r_context;

enum
{
	B2R_OK = 0,
	B2R_E_INVALID = -1,     /* bad argument */
	B2R_E_CUDA = -2,        /* CUDA runtime error (text in b2r_last_error) */
	B2R_E_NO_DEVICE = -3,   /* no usable CUDA device */
	B2R_E_UNSUPPORTED = -4, /* e.g. TriangleStrip (broken on the reference GPU path, CUDARenderer.cu:733-753) */
	B2R_E_STATE = -5        /* call order violated (e.g. read before any render) */
};

/* CullingMode.h, SampleState.h, VisualisationState.h, PrimitiveTopology.h - same numeric values */
enum { B2R_CULL_BACKFACE = 0, B2R_CULL_FRONTFACE = 1, B2R_CULL_NONE = 2 };
enum { B2R_SAMPLE_POINT = 0, B2R_SAMPLE_LINEAR = 1 };
enum { B2R_VIS_PBR = 0, B2R_VIS_DEPTH = 1, B2R_VIS_NORMAL = 2 };
enum { B2R_TOPOLOGY_TRIANGLE_LIST = 0, B2R_TOPOLOGY_TRIANGLE_STRIP = 1 };

/* options for b2r_set_option */
enum
{
	/* coverage rule for fragments lying EXACTLY on an edge (edge function == 0):
	 * 1 (default) = top-left rule, 0 = the reference's rule "all three >= 0" (CUDAROPs.cu:144). */
	B2R_OPT_FILL_RULE_TOP_LEFT = 1,
	/* sort-first sharding: this context rasterises only tiles t with t % B2R_OPT_TILE_OWNER_COUNT
	 * == B2R_OPT_TILE_OWNER_RANK (t = tx + ty*tiles_x).  Defaults 0 / 1 (owns every tile). */
	B2R_OPT_TILE_OWNER_RANK = 2,
	B2R_OPT_TILE_OWNER_COUNT = 3,
	/* also write the parity surfaces float z / winning primitive id per pixel (costs 8 B/px). Default 0. */
	B2R_OPT_DEBUG_SURFACES = 4,
	/* record cudaEvents between stages so b2r_get_stage_ms works. Default 0. */
	B2R_OPT_STAGE_TIMERS = 5,
	/* 1: every tile list is put into ascending primitive-id order INSIDE the frame (by the raster kernel; longer lists by a
	 * sorter kernel), so the list buffers themselves are bit-reproducible.  0 (default): the lists stay in the order the
	 * scatter atomics were served - the frame does not depend on it (the depth resolve is an order-independent min over
	 * (depth, primitive id)) - and b2r_read_bins puts them into the same canonical order when they are read.  Measured on
	 * BASELINE config 4 (64 full-screen quads: 128-entry lists whose runs interleave) the in-frame sort is a third of
	 * the raster kernel's time. */
	B2R_OPT_SORT_BINS = 6,
	/* count the fragments that pass the inside and z-range tests (the F of the fill/ROP roofline
	 * contract) into counter [6]; an instrumented variant of the raster kernel. Default 0. */
	B2R_OPT_COUNT_FRAGMENTS = 7,
	/* sort-first: write the colours of the owned tiles tile-major (owned tile k, sub-tile s at
	 * (k*S + s)*256 pixels) instead of row-major, so one rank's share is one contiguous buffer. Default 0. */
	B2R_OPT_PACKED_OUTPUT = 8,
	/* run the vertex stage of a frame on a second stream so that it overlaps binning / raster of the previous
	 * frame when frames are enqueued back to back (transformed positions and per-mesh constants are
	 * double-buffered).  Ignored while B2R_OPT_STAGE_TIMERS is on.  Default 1. */
	B2R_OPT_OVERLAP_VERTEX = 10,
	/* pipeline whole frames: run the complete geometry front end of a frame (constants, vertex, setup, binning,
	 * sort) on the second stream, so that it overlaps the raster kernel of the PREVIOUS frame when frames are
	 * enqueued back to back; every buffer the front end writes is double-buffered by frame parity.  Results are
	 * unchanged.  Needs B2R_OPT_OVERLAP_VERTEX; ignored while B2R_OPT_STAGE_TIMERS is on.  With the sharded front
	 * end (b2r_shard_*) a frame's front end additionally waits for the previous frame's "lists ready" barrier
	 * before it stores into the peers' list sets.  Default 1. */
	B2R_OPT_OVERLAP_FRONT = 11,
	/* frame ring, the alternative to the two options above (EXPERIMENTAL in round 1: written after the round's GPU budget
	 * was spent, default 0 = off; also B2R_FRAME_RING=D in the environment).  Value D = 2..4: frame f runs from its
	 * constants to its raster kernel on ring stream f % D and owns slot f % D of everything a frame writes, so the
	 * geometry front ends of up to D consecutive frames are in flight at once; only the raster kernels are chained in
	 * frame order.  Meant for the sharded front end (b2r_shard_*), where a rank's front end is a chain of short
	 * latency-bound kernels: the arena then holds D list sets, every rank announces "frame done" to all ranks, and a
	 * rank waits for the peers' announcement of frame f-D before it stores the lists of frame f into their set f % D.
	 * Set it between frames and before b2r_shard_prepare.  Results are unchanged. */
	B2R_OPT_FRAME_RING = 12,
	/* sharded front end + frame ring only (EXPERIMENTAL, default 0; also B2R_SHARD_SLICED_VERTEX=1): the vertex stage
	 * of a rank transforms only the vertices its own primitive slice refers to (1/n of the replicated work) and the
	 * raster kernel transforms the three vertices of a list entry itself, with the vertex stage's own function on the
	 * resident model-space positions - same bits.  b2r_read_overtex is not available in this mode. */
	B2R_OPT_SHARD_SLICED_VERTEX = 13,
	/* sharded front end: every rank keeps the colours of the tiles it owns in its OWN colour surface instead of
	 * storing them into rank 0's over NVLink (no "colours landed" flags either).  For consumers on the host
	 * (b2r_present_tiles_async: every rank delivers its tiles over its own PCIe link); the frame is then never
	 * composed on a device.  Set it between frames, after b2r_shard_bind.  Default 0. */
	B2R_OPT_SHARD_LOCAL_COLOUR = 14,
	/* Direct path for small triangles: a triangle with at most this many candidate pixels (its conservative pixel box
	 * inside the bins of its rectangle - the pixels the reference's per-bin test, CUDARenderer.cu:1116-1127, can pass)
	 * is resolved by the setup kernel itself with 64-bit atomicMin operations on a per-pixel (depth, primitive id) key
	 * buffer instead of going through the bin lists; the raster kernel merges the key buffer into its tiles.  Same
	 * fragments, same keys, same order-independent min: the frame is identical bit for bit.  0 switches the path off
	 * (every visible triangle is binned).  b2r_read_bins always returns the complete bins (it re-runs the newest frame
	 * with the path off when needed).  Sharded front end: the ranks exchange the touched 16x16 key blocks instead of list entries (needs 16x16 bins; the switch is fixed at b2r_shard_prepare and must agree on all ranks).  Default 16; set it between frames. */
	B2R_OPT_DIRECT_MAX_AREA = 15
};

/* stage indices for b2r_get_stage_ms */
enum
{
	B2R_STAGE_VERTEX = 0, B2R_STAGE_SETUP = 1, B2R_STAGE_BIN_SCAN = 2, B2R_STAGE_BIN_SCATTER = 3,
	B2R_STAGE_BIN_SORT = 4, B2R_STAGE_RASTER_SHADE = 5, B2R_STAGE_TOTAL = 6,
	/* sharded front end only (0 otherwise): EXCHANGE = pulling the peers' lists over NVLink,
	 * PEER_WAIT = time inside the two flag barriers (waiting for the slowest rank's lists / colours) */
	B2R_STAGE_EXCHANGE = 7, B2R_STAGE_PEER_WAIT = 8, B2R_NUM_STAGES = 9
};

/* error text of the last failing call on this context (NULL context: last b2r_create failure) */
const char* b2r_last_error(const b2r_context* ctx);

/* CUDARenderer::CUDARenderer(const WindowHelper&) + AllocateCUDADeviceBuffers
 * (CUDARenderer.cu:34-46, 414-475; CUDAWindowHelper.h:9-31).  Bins default to 16x16 tiles. */
int b2r_create(int device, uint32_t width, uint32_t height, b2r_context** out);
/* CUDARenderer::~CUDARenderer / FreeCUDADeviceBuffers (CUDARenderer.cu:49-53, 515-525) */
void b2r_destroy(b2r_context* ctx);

/* CUDARenderer::SetupBins(numBins, binDim, binQueueMaxSize) (CUDARenderer.cu:200-204).
 * queue_max_size is accepted for source compatibility and ignored: bin lists are exact-sized
 * (count -> scan -> scatter), so the reference's 256-triangle batching disappears. */
int b2r_setup_bins(b2r_context* ctx, uint32_t num_bins_x, uint32_t num_bins_y,
                   uint32_t bin_w, uint32_t bin_h, uint32_t queue_max_size);

/* CUDATexture::Create (CUDATexture.cpp:24-72) minus the PNG decode: pixels are RGBA8,
 * byte 0 = R (SDL_image order; RGBRaw.h:34-38), tightly packed w*h words.
 * Ids are dense insertion indices like CUDATextureManager::AddCUDATexture (CUDATextureManager.cpp:29-37). */
int b2r_texture_create_rgba8(b2r_context* ctx, const uint32_t* pixels, uint32_t w, uint32_t h, int* out_tex_id);
/* CUDATexture::Destroy (CUDATexture.cpp:74-82); the id is not reused */
int b2r_texture_destroy(b2r_context* ctx, int tex_id);

/* CUDAMesh::CUDAMesh + Allocate (CUDAMesh.cpp:4-20, 42-59) and CUDARenderer::UpdateCUDAMeshTextures
 * (CUDARenderer.cu:543-573).  ivertex56 = n_vertices records of {p3, uv2, n3, tan3, c3} floats
 * (Vertex.h:113-145).  world = the matrix the reference's CUDAMesh would hold (the veneer applies
 * the double translation of CUDAMesh.cpp:16-18).  tex_ids = {diffuse, normal, specular, gloss}, -1 = none.
 * Meshes are drawn in upload order (CUDARenderer.cu:237). */
int b2r_mesh_upload(b2r_context* ctx, const void* ivertex56, uint32_t n_vertices,
                    const uint32_t* indices, uint32_t n_indices, int topology,
                    const float world[16], const int tex_ids[4], int* out_mesh_id);
/* same vertex/index buffers as src_mesh_id under another world matrix / texture set
 * (instancing instead of N copies) */
int b2r_mesh_instance(b2r_context* ctx, int src_mesh_id, const float world[16], const int tex_ids[4], int* out_mesh_id);
/* per-frame world refresh (the reference snapshots it at LoadScene, CUDAMesh.cpp:16) */
int b2r_mesh_set_world(b2r_context* ctx, int mesh_id, const float world[16]);
/* CUDARenderer::FreeAllCUDAMeshBuffers (CUDARenderer.cu:489-496) */
int b2r_mesh_clear_all(b2r_context* ctx);

int b2r_set_option(b2r_context* ctx, int option, int value);
/* current value of an option (the defaults of some options come from the environment, see above; the ranks of a
 * sharded group must agree on B2R_OPT_FRAME_RING and B2R_OPT_SHARD_SLICED_VERTEX) */
int b2r_get_option(b2r_context* ctx, int option, int* value);

/* CUDARenderer::EnterValidRenderingState/Clear + Render (CUDARenderer.cu:374-381, 1626-1659, 207-354).
 * cam_pos / cam_fwd = Camera::GetPosition/GetForward, view_proj = projection * Inverse(lookAt)
 * (CUDARenderer.cu:215-219).  clear_rgb as CUDARenderer::Clear's colour (default .25,.25,.25,
 * CUDARenderer.h:135).  Asynchronous: work is enqueued on the context's stream. */
int b2r_render(b2r_context* ctx, const float cam_pos[3], const float cam_fwd[3], const float view_proj[16],
               int cull_mode, int sample_state, int vis_state, const float clear_rgb[3]);
/* the cudaDeviceSynchronize that ends CUDARenderer::Render (CUDARenderer.cu:339) */
int b2r_sync(b2r_context* ctx);

/* CUDARenderer::Present's device->host copy (CUDARenderer.cu:386-388): width*height words 0x00RRGGBB.
 * Synchronous.  Full-speed if `out` is pinned memory. */
int b2r_read_framebuffer(b2r_context* ctx, uint32_t* out);
/* Present without the stall: the colour surface of the newest frame is copied to
 * `host_pixels` (pinned memory) on a second stream; the call returns immediately and the next frame
 * renders into the other of two colour surfaces, so the copy overlaps the next b2r_render.
 * b2r_present_wait blocks until every outstanding copy has landed (and the frames are complete). */
int b2r_present_async(b2r_context* ctx, uint32_t* host_pixels);
int b2r_present_wait(b2r_context* ctx);
/* CUDARenderer::DrawTexture(char*) / DrawTextureGlobal(char*, bool, SampleState) (CUDARenderer.h:43-45, CUDARenderer.cu:1968-2012;
 * "testing purposes" in the reference): fills the colour surface with the RGBA8 texture `pixels` (byte 0 = R) - global = 0: one
 * point-filtered texel per pixel at normalised window coordinates, written with bytes 0 and 2 swapped (TextureTestKernel);
 * global = 1: CUDATextureSampler::Sample at x / width (stretched) or x / texture width, packed like a shaded pixel
 * (DrawTextureGlobalKernel).  Synchronous; the result is read with b2r_read_framebuffer. */
int b2r_draw_texture(b2r_context* ctx, const uint32_t* pixels, uint32_t width, uint32_t height, int global, int stretched, int sample_state,
                     const float clear_rgb[3]);

/* Distributed Present (sort-first over several GPUs; CUDARenderer::Present, CUDARenderer.cu:383-395, when the
 * consumer of the frame is the host): the frame is composed in HOST memory that all ranks share - e.g. a POSIX
 * shared-memory mapping - and every rank delivers the tiles it owns over its own PCIe link, instead of all colours
 * travelling to rank 0 and leaving through one link.
 *   b2r_host_frame_register(ctx, host, bytes)   page-locks and maps the caller's memory for this context's device
 *                                               (cudaHostRegister); host = NULL, bytes = 0 releases it
 *   b2r_present_tiles_async(ctx, byte_offset)   starts copying this rank's tiles (B2R_OPT_TILE_OWNER_*; all tiles
 *                                               on a single GPU) of the newest frame into the row-major width x
 *                                               height frame at host + byte_offset; returns immediately; the next
 *                                               frame renders into the other colour surface (as b2r_present_async)
 * b2r_present_wait blocks until this rank's copies have landed; a consumer of the composed frame additionally
 * needs a barrier across the ranks.  With the sharded front end use B2R_OPT_SHARD_LOCAL_COLOUR. */
int b2r_host_frame_register(b2r_context* ctx, void* host, uint64_t bytes);
int b2r_present_tiles_async(b2r_context* ctx, uint64_t byte_offset);
/* Distributed Present of the sharded front end in its default form (frame composed in rank 0's surface over NVLink):
 * rank r copies the r-th band of tile rows of the composed frame - a contiguous range of the row-major surface - from
 * rank 0's memory (mapped on every rank by b2r_shard_bind) into the registered host frame at host + byte_offset,
 * with its own copy engine over its own PCIe link, after waiting (on the device) until every rank's colours of the
 * frame have landed on rank 0.  All ranks call it after the same b2r_render; the n bands make up the frame.
 * b2r_present_wait blocks until this rank's band has landed; a consumer needs a barrier across the ranks.  A rank's
 * next frame of the same parity waits for this rank's copy before it arrives at the frame's "lists ready" barrier,
 * so no rank overwrites the surface while any band of it is still being read.  On a context without a sharded group
 * the band is the whole frame. */
int b2r_present_band_async(b2r_context* ctx, uint64_t byte_offset);
/* The same with the band given by the caller - rows [row_begin, row_end) of the frame - for boxes whose GPUs do not reach the
 * host equally fast: the ranks' bands must tile the frame, their sizes are the caller's choice (bench.py sizes them by a
 * short calibration of every rank's device->host bandwidth with all ranks copying at once). */
int b2r_present_rows_async(b2r_context* ctx, uint64_t byte_offset, uint32_t row_begin, uint32_t row_end);
/* depth buffer as the reference keeps it (int z*2^31, INT_MAX = empty; CUDAROPs.cu:386, CUDARenderer.cu:1640) */
int b2r_read_depth(b2r_context* ctx, int32_t* out);
/* parity surfaces (need B2R_OPT_DEBUG_SURFACES=1 before the render): PixelShade.zInterpolated
 * (0 where empty, like the zero-cleared fragment buffer) and the winning global primitive id
 * (0xFFFFFFFF where empty) */
int b2r_read_z(b2r_context* ctx, float* out);
int b2r_read_winner(b2r_context* ctx, uint32_t* out);
/* OVertex buffer of one mesh, 18 floats per vertex {p4, uv2, n3, tan3, vd3, c3} (Vertex.h:280-327) */
int b2r_read_overtex(b2r_context* ctx, int mesh_id, float* out18);
/* per-primitive visibility of one mesh after cull + frustum reject (the set TA_Kernel appends,
 * CUDARenderer.cu:702-796), one byte per triangle */
int b2r_read_visible(b2r_context* ctx, int mesh_id, uint8_t* out);
/* bin lists of the last frame: counts[num_bins], exclusive offsets[num_bins], and the
 * concatenated global primitive ids (mesh order, then triangle index).  Pass ids = NULL to query
 * sizes; *total receives the number of (triangle, bin) pairs. */
int b2r_read_bins(b2r_context* ctx, uint32_t* counts, uint32_t* offsets, uint32_t* ids, uint64_t ids_capacity, uint64_t* total);

/* CUDARenderer::GetTotalNumTriangles / GetTotalNumVisibleTriangles (CUDARenderer.h:63-64) */
int b2r_get_stats(b2r_context* ctx, uint32_t* total_tris, uint32_t* visible_tris);
/* device time of each stage of the last frame rendered with B2R_OPT_STAGE_TIMERS
 * (replaces the BENCHMARK macro, CUDARenderer.cu:223-353) */
int b2r_get_stage_ms(b2r_context* ctx, float* out_ms, int n);
/* mean stage times over all timed frames since the last reset; *frames = how many */
int b2r_get_stage_ms_mean(b2r_context* ctx, float* out_ms, int n, uint64_t* frames, int reset);
/* counters of the last frame: [0] (triangle,bin) pairs, [1] kernels launched by b2r_render,
 * [2] tiles rasterised by this context, [3] number of bins, [4] tile lists left unsorted (longer
 * than the sorter handles), [5] triangles binned cooperatively, [6] fragments (B2R_OPT_COUNT_FRAGMENTS) */
int b2r_get_counters(b2r_context* ctx, uint64_t* out, int n);

/* interop for harnesses that own device memory / streams (PyTorch, NCCL):
 * run on a caller-provided cudaStream_t (NULL = the context's own stream) */
int b2r_set_stream(b2r_context* ctx, void* cuda_stream);
/* device pointers of the colour (uint32) and depth (int32) surfaces */
int b2r_device_surfaces(b2r_context* ctx, void** framebuffer, void** depth);
/* redirect the ROP write-back of later frames into caller-provided device memory
 * (e.g. a peer GPU's framebuffer mapped over NVLink); NULL restores the internal surface */
int b2r_bind_framebuffer(b2r_context* ctx, void* device_ptr);


/* ---- multi-GPU (sort-first) helpers; the reference is single-GPU (main.cpp:397-399), nothing replaced ----
 * pixels of the packed buffer a rank with this ownership writes under B2R_OPT_PACKED_OUTPUT */
int b2r_packed_size(b2r_context* ctx, int rank, int count, uint64_t* pixels);
/* on the gathering rank: scatter one rank's packed colours into a row-major framebuffer
 * (dst_framebuffer = NULL: this context's own surface); runs on the context's stream */
int b2r_unpack_tiles(b2r_context* ctx, const void* packed_device, int rank, int count, void* dst_framebuffer);
/* CUDA IPC so that ranks in other processes can write their tiles straight into the gathering
 * rank's memory over NVLink (device_ptr = NULL exports this context's framebuffer); handle = 64 bytes */
int b2r_ipc_export(b2r_context* ctx, void* device_ptr, void* handle64);
int b2r_ipc_open(b2r_context* ctx, const void* handle64, void** device_ptr);
int b2r_ipc_close(b2r_context* ctx, void* device_ptr);


/* ---- sharded front end (sort-first with the geometry work divided as well) ----
 * With plain tile ownership every rank still sets up and bins the whole mesh.  In the sharded mode rank r
 * sets up / bins / sorts only the r-th contiguous slice of the primitives - into the lists of ALL tiles -
 * and then stores (posted NVLink writes into CUDA-IPC-mapped peer memory) the heads and the sorted ids of
 * every tile's list into a region reserved for it in the memory of the tile's owner.  The owner's raster
 * kernel walks the ranks' segments in rank order from local memory; because the slices are contiguous and
 * ascending, that concatenation is exactly the list a single GPU builds.  Every rank stores its colours
 * straight into rank 0's surface.  What crosses the link per frame is 4 bytes per (triangle, tile) pair
 * plus the colours; there is no collective.  Two flag barriers per frame (one small kernel each) order
 * "lists ready" and "colours landed"; the bin lists and rank 0's colour surface are double-buffered by
 * frame parity so that no other synchronisation is needed.
 *
 * Protocol (same scene loaded on every rank; one context per process/GPU; 2..8 ranks):
 *   1. every rank: b2r_shard_prepare(ctx, rank, count, list_capacity, arena_handle, colour_handle)
 *      list_capacity = entries of one rank's bin-list buffer (0 = default), identical on every rank
 *   2. all-gather the 64-byte arena handles; broadcast rank 0's colour handle
 *   3. every rank: b2r_shard_bind(ctx, handles[count][64], colour_handle_of_rank_0); host barrier
 *   4. every rank calls b2r_render with the same arguments the same number of times; the composed frame is
 *      in rank 0's surface (b2r_read_framebuffer / b2r_present_async on rank 0).  b2r_read_bins returns the
 *      calling rank's own segment lists.
 *   5. host barrier; b2r_shard_release on every rank
 * An undersized list_capacity is reported by b2r_sync (B2R_E_STATE) and the frame is void on every rank;
 * a peer that does not reach a barrier within 4 s is reported the same way instead of hanging the GPU. */
int b2r_shard_prepare(b2r_context* ctx, int rank, int count, uint64_t list_capacity, void* arena_handle64, void* colour_handle64);
int b2r_shard_bind(b2r_context* ctx, const void* arena_handles, const void* colour_handle_rank0);
int b2r_shard_release(b2r_context* ctx);

#ifdef __cplusplus
}
#endif
#endif
