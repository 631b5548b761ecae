// TEST INFRASTRUCTURE -- the 11 SDL2 / SDL2_image symbols the reference's render-path
// translation units leave undefined at link, as headless no-ops.
// Ours, not reference code.  Layout of SDL_Surface / SDL_PixelFormat comes from the
// vendored SDL 2.0.9 headers the reference itself compiles against.
//
//   CUDARenderer.cu:377-394  SDL_LockSurface SDL_UnlockSurface SDL_UpperBlit SDL_UpdateWindowSurface
//   CUDATexture.cpp:28,69    IMG_Load SDL_FreeSurface
//   EventManager.cpp         SDL_PollEvent SDL_GetKeyboardState SDL_GetRelativeMouseMode
//                            SDL_SetRelativeMouseMode SDL_GetRelativeMouseState
#include <SDL.h>
#include <SDL_image.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

namespace {
struct MemTexture { unsigned w, h; const unsigned char* rgba; };
std::map<std::string, MemTexture>& Table() { static std::map<std::string, MemTexture> t; return t; }
}

// registered by oracle/ref_driver.cu: IMG_Load("mem:<i>") then yields that RGBA8 image
extern "C" void b2r_stub_register_texture(const char* name, unsigned w, unsigned h, const unsigned char* rgba)
{
	Table()[name] = MemTexture{ w, h, rgba };
}

extern "C" {

SDL_Surface* IMG_Load(const char* file)
{
	auto it = Table().find(file ? file : "");
	if (it == Table().end())
		return nullptr;
	const MemTexture& t = it->second;
	SDL_Surface* s = (SDL_Surface*)calloc(1, sizeof(SDL_Surface));
	SDL_PixelFormat* f = (SDL_PixelFormat*)calloc(1, sizeof(SDL_PixelFormat));
	f->BitsPerPixel = 32;
	f->BytesPerPixel = 4;
	s->format = f;
	s->w = (int)t.w;
	s->h = (int)t.h;
	s->pitch = (int)t.w * 4;
	s->pixels = malloc(size_t(t.w) * t.h * 4);
	memcpy(s->pixels, t.rgba, size_t(t.w) * t.h * 4);
	return s;
}

void SDL_FreeSurface(SDL_Surface* s)
{
	if (!s) return;
	free(s->pixels);
	free(s->format);
	free(s);
}

int SDL_LockSurface(SDL_Surface*) { return 0; }
void SDL_UnlockSurface(SDL_Surface*) {}
int SDL_UpperBlit(SDL_Surface*, const SDL_Rect*, SDL_Surface*, SDL_Rect*) { return 0; }
int SDL_UpdateWindowSurface(SDL_Window*) { return 0; }
int SDL_PollEvent(SDL_Event*) { return 0; }
const Uint8* SDL_GetKeyboardState(int* numkeys)
{
	static Uint8 keys[SDL_NUM_SCANCODES];
	if (numkeys) *numkeys = SDL_NUM_SCANCODES;
	return keys;
}
SDL_bool SDL_GetRelativeMouseMode(void) { return SDL_FALSE; }
int SDL_SetRelativeMouseMode(SDL_bool) { return 0; }
Uint32 SDL_GetRelativeMouseState(int* x, int* y) { if (x) *x = 0; if (y) *y = 0; return 0; }

}
