/* stb_ref_shim.c -- test infrastructure: exposes the reference's own PNG decode call
 * (cpp/deplex/src/deplex/utils/depth_image.cpp:32: stbi_load_16(path, &w, &h, &channels, STBI_grey)) from the vendored
 * stb_image.h, compiled where it lies under the reference tree (oracle/Makefile, target `ref`), so that the zlib-based
 * reader of deplex_b200/cpp/src/depth_image.cpp can be compared with it file by file.  No reference source is copied. */
#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_PNG
#include "stb_image.h"

#include <string.h>

/* returns 0 on failure; otherwise writes up to `capacity` samples and the image size */
int stb_ref_load16_grey(const char* path, unsigned short* out, long capacity, int* width, int* height) {
  int w = 0, h = 0, ch = 0;
  unsigned short* data = stbi_load_16(path, &w, &h, &ch, STBI_grey);
  if (!data) return 0;
  *width = w;
  *height = h;
  if ((long)w * h <= capacity) memcpy(out, data, sizeof(unsigned short) * (size_t)w * h);
  stbi_image_free(data);
  return 1;
}
