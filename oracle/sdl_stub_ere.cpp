// TEST INFRASTRUCTURE -- the SDL2 symbols the reference's CPU renderer (ERenderer.cpp) needs on top of
// oracle/sdl_stub.cpp, as headless stand-ins.  Ours, not reference code; struct layouts come from the vendored
// SDL 2.0.9 headers the reference compiles against.
//
//   ERenderer.cpp:22-34    SDL_GetWindowSurface SDL_GetWindowSize SDL_CreateRGBSurface
//   ERenderer.cpp:422,490  SDL_MapRGB   (back buffer: SDL_CreateRGBSurface(0,w,h,32,0,0,0,0) = XRGB8888)
//   ERenderer.cpp:174-177  SDL_SaveBMP -> SDL_SaveBMP_RW + SDL_RWFromFile
//   Texture.h:41,101,119   SDL_GetRGB   (IMG_Load of an RGBA PNG on a little-endian host = ABGR8888: R in the low byte)
#include <SDL.h>
#include <cstdlib>
#include <cstring>

namespace { int g_w = 0, g_h = 0; }

// set by oracle/ere_ref_driver.cpp before it constructs Elite::Renderer: the size of the (non-existent) window
extern "C" void b2r_stub_set_window_size(int w, int h) { g_w = w; g_h = h; }

extern "C" {

SDL_Surface* SDL_GetWindowSurface(SDL_Window*) { return nullptr; }     // front buffer: only ever the target of the stubbed blit
void SDL_GetWindowSize(SDL_Window*, int* w, int* h) { if (w) *w = g_w; if (h) *h = g_h; }

SDL_Surface* SDL_CreateRGBSurface(Uint32, int width, int height, int depth, Uint32, Uint32, Uint32, Uint32)
{
	SDL_Surface* s = (SDL_Surface*)calloc(1, sizeof(SDL_Surface));
	SDL_PixelFormat* f = (SDL_PixelFormat*)calloc(1, sizeof(SDL_PixelFormat));
	f->BitsPerPixel = (Uint8)depth;
	f->BytesPerPixel = (Uint8)(depth / 8);
	f->Rmask = 0x00FF0000u; f->Gmask = 0x0000FF00u; f->Bmask = 0x000000FFu;
	f->Rshift = 16; f->Gshift = 8; f->Bshift = 0;
	s->format = f;
	s->w = width; s->h = height; s->pitch = width * (depth / 8);
	s->pixels = calloc((size_t)width * height, (size_t)(depth / 8));
	return s;
}

Uint32 SDL_MapRGB(const SDL_PixelFormat*, Uint8 r, Uint8 g, Uint8 b) { return ((Uint32)r << 16) | ((Uint32)g << 8) | (Uint32)b; }

void SDL_GetRGB(Uint32 pixel, const SDL_PixelFormat*, Uint8* r, Uint8* g, Uint8* b)
{
	*r = (Uint8)(pixel & 0xFFu); *g = (Uint8)((pixel >> 8) & 0xFFu); *b = (Uint8)((pixel >> 16) & 0xFFu);
}

SDL_RWops* SDL_RWFromFile(const char*, const char*) { return nullptr; }
int SDL_SaveBMP_RW(SDL_Surface*, SDL_RWops*, int) { return -1; }

}
