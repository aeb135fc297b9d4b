/* iba_oracle.h — flat C interface of the CPU restatement of PtzIncrementalOptimizer (oracle/iba_oracle.cpp).
 * TEST INFRASTRUCTURE ONLY: nothing under ptz-calib_b200/ or include/ may include, link or call this. */
#ifndef PTZ_IBA_ORACLE_H
#define PTZ_IBA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* events of one Solve, in the order the driver takes its decisions: triples (kind, a, b) */
enum {
  ORC_IBA_SEED_PAIR = 1,   /* a, b = image ids FindInitialImagePair returned (ptz_incremental_optimizer.cc:142-177) */
  ORC_IBA_INIT_RESULT = 2, /* a = RegisterInitialImagePair succeeded (:352-375), b = LM iterations of its two-view BA */
  ORC_IBA_GLOBAL_BA = 3,   /* a = AdjustGlobalBundle succeeded (:421-439), b = registered images at that point */
  ORC_IBA_REGISTER = 4,    /* a = image id handed to RegisterNextImage (:377-419), b = registered neighbour it succeeded from, or -1 */
  ORC_IBA_UNREGISTER = 5,  /* a = image erased again after a failed global BA (:95-98) */
  ORC_IBA_NEXT_LIST = 6    /* a = length of the list FindNextImages returned (:249-290), b = its first image id (-1 when empty) */
};

typedef struct orc_iba_input {
  int32_t num_images;
  const int32_t* img_w;        /* [num_images] features_[i].img_size */
  const int32_t* img_h;
  const int64_t* kp_offset;    /* [num_images+1] */
  const float* kp_uv;          /* [.. * 2] */
  int32_t num_pairs;           /* matches_info_.size() */
  const int32_t* pair_src;
  const int32_t* pair_dst;
  const int64_t* match_offset; /* [num_pairs+1] */
  const int32_t* query_idx;
  const int32_t* train_idx;
  const double* H;             /* [num_pairs*9] row-major */
  const uint8_t* has_H;        /* !H.empty() */
  const double* confidence;
  const double* cams21;        /* [num_images*21] krt21 the driver starts from */
  int32_t max_iter;
  int32_t num_seeds;           /* SetSeedImageId, or 0 */
  const int64_t* seed_ids;
} orc_iba_input;

typedef struct orc_iba_output {
  int32_t ok;                  /* Solve's return value */
  int32_t num_registered;
  uint8_t* registered;         /* [num_images] */
  double* cams21;              /* [num_images*21] */
  int64_t* events;             /* [cap_events*3] */
  int32_t cap_events;
  int32_t num_events;          /* events produced (may exceed cap_events: the log is then truncated) */
  double last_reproj_error;    /* final_reproj_error_all of the last global BA */
} orc_iba_output;

int orc_iba_solve(const orc_iba_input* in, orc_iba_output* out);

#ifdef __cplusplus
}
#endif
#endif
