// TEST INFRASTRUCTURE ONLY (oracle/): PPM-backed implementation of the three SDL entry points the
// reference texture loader calls (reference VerStarting/texture.cc:68,64,83).  Surfaces are always
// produced as tightly packed RGBA32, so the reference never needs to convert.
#include <SDL2/SDL_image.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

// Reads one whitespace/comment separated decimal token of a PPM header.
bool ReadHeaderInt(FILE *f, int *out) {
  int c = fgetc(f);
  for (;;) {
    while (c == ' ' || c == '\t' || c == '\r' || c == '\n') c = fgetc(f);
    if (c == '#') {
      while (c != '\n' && c != EOF) c = fgetc(f);
      continue;
    }
    break;
  }
  if (c < '0' || c > '9') return false;
  long v = 0;
  while (c >= '0' && c <= '9') {
    v = v * 10 + (c - '0');
    if (v > 1000000) return false;
    c = fgetc(f);
  }
  // The single whitespace byte after the token has been consumed (c), as the PPM grammar wants.
  *out = (int)v;
  return true;
}

}  // namespace

extern "C" SDL_Surface *IMG_Load(const char *file) {
  FILE *f = fopen(file, "rb");
  if (f == nullptr) return nullptr;
  char magic[2] = {0, 0};
  int w = 0, h = 0, maxval = 0;
  if (fread(magic, 1, 2, f) != 2 || magic[0] != 'P' || magic[1] != '6' ||
      !ReadHeaderInt(f, &w) || !ReadHeaderInt(f, &h) || !ReadHeaderInt(f, &maxval) ||
      maxval != 255 || w <= 0 || h <= 0) {
    fclose(f);
    return nullptr;
  }
  size_t n = (size_t)w * (size_t)h;
  unsigned char *rgb = (unsigned char *)malloc(n * 3);
  if (rgb == nullptr || fread(rgb, 3, n, f) != n) {
    free(rgb);
    fclose(f);
    return nullptr;
  }
  fclose(f);

  SDL_Surface *s = (SDL_Surface *)calloc(1, sizeof(SDL_Surface));
  s->format = (SDL_PixelFormat *)calloc(1, sizeof(SDL_PixelFormat));
  s->format->format = SDL_PIXELFORMAT_RGBA32;
  s->w = w;
  s->h = h;
  s->pitch = w * 4;
  unsigned char *px = (unsigned char *)malloc(n * 4);
  for (size_t i = 0; i < n; i++) {
    px[i * 4 + 0] = rgb[i * 3 + 0];
    px[i * 4 + 1] = rgb[i * 3 + 1];
    px[i * 4 + 2] = rgb[i * 3 + 2];
    px[i * 4 + 3] = 255;
  }
  free(rgb);
  s->pixels = px;
  return s;
}

extern "C" void SDL_FreeSurface(SDL_Surface *surface) {
  if (surface == nullptr) return;
  free(surface->pixels);
  free(surface->format);
  free(surface);
}

extern "C" SDL_Surface *SDL_ConvertSurfaceFormat(SDL_Surface *src, uint32_t, uint32_t) {
  // IMG_Load above only ever produces RGBA32, so the reference never gets here; fail loudly if it does.
  (void)src;
  fprintf(stderr, "sdl_stub: unexpected SDL_ConvertSurfaceFormat call\n");
  return nullptr;
}
