/* dbde_util.h -- drop-in C++ interface of the B200-native DBDE codec.
 *
 * Same sixteen free functions, same structs, same C++ linkage (hence the same mangled symbols)
 * as the reference header /root/reference/dbde_util.h:8-52, so code written against the
 * reference -- including its own dbde_util_test.cpp -- links against libdbde_b200.so unchanged.
 * Unlike the reference header this one is self-contained (it pulls in the integer and stdio
 * types it uses).  Frame encode/decode run on the GPU through include/dbde_b200.h; header
 * marshalling and the file walker are host code.
 *
 * Conventions kept from the reference:
 *   - the caller owns every buffer; `target` must hold 32 + 66*ceil(W/8)*ceil(H/8) bytes
 *   - no exceptions / errno: dbde_unpack_image returns 0 on a malformed block and leaves the
 *     image untouched; header parsers flag errors with u64s == (uint32_t)-1
 *   - calls are synchronous; distinct threads may call concurrently on distinct buffers
 */
#ifndef DBDE_UTIL
#define DBDE_UTIL

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

/* in-memory headers (on disk: 28 and 20 packed little-endian bytes) */
struct video_header {
    uint32_t u64s;      /* 3 when valid */
    uint64_t height;
    uint64_t width;
    double frame_hz;
};

struct frame_header {
    uint32_t u64s;      /* 2 when valid */
    uint64_t index;
    uint64_t elapsed_ns;   /* travels on disk as an IEEE-754 double (dbde_util.cpp:186,334) */
};

/* streaming reader state (reference dbde_util.h:39-48) */
struct dbde_file_walker {
    FILE *fptr;
    int32_t frames;
    size_t i;
    size_t n;
    size_t N;
    int32_t width;
    int32_t height;
    uint8_t *buffer;
};

/* ---- encode (reference dbde_util.h:21-28) ---- */
uint32_t dbde_pack_8x8(uint8_t *image, int stride, uint8_t *target);
uint32_t dbde_pack_8x8_partial(uint8_t *image, int stride, int rightmargin, int downmargin, uint8_t *target);
size_t dbde_pack_image(uint8_t *image, int W, int H, uint8_t *target);
size_t dbde_pack_frame_header(frame_header fh, uint8_t *target);
size_t dbde_pack_frame(uint64_t index, uint8_t *image, int W, int H, uint8_t *target);
size_t dbde_pack_video_header(video_header vh, uint8_t *target);

/* ---- decode (reference dbde_util.h:30-37) ---- */
void dbde_unpack_8x8(uint8_t depth, uint8_t minval, uint8_t *packed, size_t stride, uint8_t *image);
void dbde_unpack_8x8_partial(uint8_t depth, uint8_t minval, uint8_t *packed, size_t stride, int rightmargin,
                             int downmargin, uint8_t *image);
size_t dbde_unpack_image(uint8_t *packed, int W, int H, uint8_t *image);
frame_header dbde_unpack_frame_header(uint8_t **packed);
frame_header dbde_unpack_frame(uint8_t **packed, int W, int H, uint8_t *image);
video_header dbde_unpack_video_header(uint8_t **packed);

/* ---- file walker (reference dbde_util.h:50-52) ---- */
dbde_file_walker dbde_start_file_walk(const char *name, int frames_buffered, video_header *vh);
bool dbde_walk_a_file(dbde_file_walker *walker, frame_header *fh, uint8_t *image);
void dbde_end_file_walk(dbde_file_walker *walker);

#endif
