// aligner.hpp -- device-resident coarse-to-fine aligner for `batch` independent frame pairs.
#pragma once
#include <vector>
#include "ctx.hpp"

namespace rgbid {

enum MapId {
  MAP_W_KF = 0, MAP_I_KF, MAP_GWX, MAP_GWY, MAP_GIX, MAP_GIY, MAP_W_CUR, MAP_I_CUR,
  MAP_CGWX, MAP_CGWY, MAP_CGIX, MAP_CGIY,  // covariance-only (bilateral-filtered) gradients, tracker mode
  MAP_WF, MAP_IF,                          // bilateral-filtered keyframe pyramid (scratch), tracker mode
  MAP_W_WARP, MAP_I_WARP,                  // WARP_ORDER = warpFirst: per-iteration pyramid of the warped current frame
  MAP_COUNT
};

struct LevelGeom {
  int rows, cols;
  size_t pitch, sstride;
  int kept_rows, kept_cols, sample_stride;
};

}  // namespace rgbid

struct rgbid_aligner {
  rgbid_ctx* ctx;
  rgbid_align_config cfg;
  int niters;        // Gauss-Newton iterations per pair
  int trace_stride;  // niters + 1 (covariance pass)
  rgbid::LevelGeom geom[RGBID_MAX_LEVELS];
  char* d_arena;
  size_t arena_bytes;
  rgbid::ImgB maps[rgbid::MAP_COUNT][RGBID_MAX_LEVELS];
  rgbid::GnState* d_states;
  rgbid::ScaleState* d_scales;
  double* d_partials;
  int partial_blocks;
  unsigned int* d_counters;
  rgbid_iter_trace* d_trace;
  double* d_init;  // batch x (9 + 3)
  uint16_t* d_depth_raw;
  uint8_t* d_rgb_raw;
  int* d_active;   // per-stream predicate for keyframe updates (tracker)
  // pinned host mirrors
  rgbid::GnState* h_states;
  int* d_trace_flag;   // device: non-zero = the solver tail writes the per-iteration trace (rgbid_aligner_set_trace)
  int trace_enabled;   // host shadow of it
  rgbid_iter_trace* h_trace;
  double* h_init;
  int* h_active;
  // CUDA graph of the whole schedule
  bool use_graph;
  bool use_pdl;  // programmatic dependent launch between the Gauss-Newton kernels (RGBID_NO_PDL=1 disables)
  cudaGraphExec_t gn_exec[3];       // indexed by ALIGN_PART_*
  long long gn_graph_launches[3];
  int image_filtering;
  // second stream + fork / join events: the schedule is issued as two chains of frame-pair groups
  cudaStream_t side_stream;
  cudaEvent_t ev_fork, ev_join;
  // texture objects over the current-frame pyramid: [level][0: W point | 1: I linear][batch]
  bool use_tex;
  cudaTextureObject_t* h_tex;
  cudaTextureObject_t* d_tex;

  rgbid::ImgB view(int which, int level, int index) const
  {
    rgbid::ImgB m = maps[which][level];
    m.p = (float*)((char*)m.p + (size_t)index * m.sstride);
    return m;
  }
};

namespace rgbid {

// Build pyramid + gradients (+ filtered gradients in tracker mode) of the keyframe maps of `batch` streams
// starting at `first`; active (device, may be null) predicates streams.
void aligner_keyframe_derivatives(rgbid_aligner* al, int first, int batch, const int* active, bool pyramid_from_l0);
void aligner_current_pyramid(rgbid_aligner* al, int first, int batch);
void aligner_copy_current_to_keyframe(rgbid_aligner* al, int first, int batch, const int* active);
enum { ALIGN_PART_ALL = 0, ALIGN_PART_ITERATIONS = 1, ALIGN_PART_COV = 2 };
void aligner_record_schedule(rgbid_aligner* al, int part = ALIGN_PART_ALL);
int aligner_enqueue_device_init(rgbid_aligner* al);
// tracker mode: the Gauss-Newton iterations and the covariance pass as two separately launched graphs
bool aligner_can_split(const rgbid_aligner* al);
int aligner_enqueue_part(rgbid_aligner* al, int part);

}  // namespace rgbid
