// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for <SDL2/SDL_image.h>.
//
// The reference's texture.cc (reference VerStarting/texture.cc:1,60-109) decodes image files through
// SDL2_image, which is not installed in this image.  So that the UNMODIFIED reference texture.cc
// (GetColorAt and LoadFromFile, including the px/255.0 conversion at texture.cc:100-104) can be
// compiled into oracle/_ref, this header declares only the handful of SDL names that file touches
// and sdl_stub.cc implements them for binary PPM (P6, maxval 255) files.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDL_PIXELFORMAT_RGBA32 0x16762004u

typedef struct SDL_PixelFormat {
  uint32_t format;
} SDL_PixelFormat;

typedef struct SDL_Surface {
  uint32_t flags;
  SDL_PixelFormat *format;
  int w, h;
  int pitch;
  void *pixels;
} SDL_Surface;

SDL_Surface *IMG_Load(const char *file);
void SDL_FreeSurface(SDL_Surface *surface);
SDL_Surface *SDL_ConvertSurfaceFormat(SDL_Surface *src, uint32_t pixel_format, uint32_t flags);

#ifdef __cplusplus
}
#endif
