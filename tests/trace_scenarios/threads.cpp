// Scenario for tests/test_cpu_trace.py (ThreadSanitizer build of the tracing library): the C ABI is "thread-compatible per
// handle"  - independent contexts may be driven from different host threads at the same time.  Four
// threads each create a context, upload a mesh and a texture, render frames in flight, read back, change options and
// destroy it, concurrently; a fifth hammers the error path of b2r_create.  TSan reports any shared mutable state in the
// product's host code.
#include "b2r.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

static std::atomic<int> g_failures{ 0 };
#define CHECK(call) do { const int rc_ = (call); if (rc_ != B2R_OK) { fprintf(stderr, "%s -> %d\n", #call, rc_); ++g_failures; return; } } while (0)

static void worker(int id)
{
	b2r_context* c = nullptr;
	CHECK(b2r_create(0, 320, 240, &c));
	std::vector<float> v(3 * 14 * 64, 0.f);
	std::vector<uint32_t> idx(3 * 64);
	for (size_t i = 0; i < idx.size(); ++i) { idx[i] = (uint32_t)i; v[i * 14] = float(i % 7) - 3.f; v[i * 14 + 1] = float(i % 5) - 2.f; v[i * 14 + 2] = -float(id + 1); }
	std::vector<uint32_t> px(64 * 64, 0xFF00FF00u);
	int tex = -1, mesh = -1;
	CHECK(b2r_texture_create_rgba8(c, px.data(), 64, 64, &tex));
	const float world[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	const int texIds[4] = { tex, -1, -1, -1 };
	CHECK(b2r_mesh_upload(c, v.data(), (uint32_t)(v.size() / 14), idx.data(), (uint32_t)idx.size(), B2R_TOPOLOGY_TRIANGLE_LIST, world, texIds, &mesh));
	const float pos[3] = { 0, 0, 5 }, fwd[3] = { 0, 0, 1 };
	std::vector<uint32_t> fb(320 * 240);
	for (int round = 0; round < 6; ++round)
	{
		if (round == 3) CHECK(b2r_set_option(c, B2R_OPT_FRAME_RING, 2 + id % 3));
		for (int f = 0; f < 9; ++f) CHECK(b2r_render(c, pos, fwd, world, 2, f & 1, 0, nullptr));
		CHECK(b2r_read_framebuffer(c, fb.data()));
		uint32_t total = 0, visible = 0;
		CHECK(b2r_get_stats(c, &total, &visible));
		if (b2r_read_framebuffer(c, nullptr) == B2R_OK || !b2r_last_error(c)[0]) { ++g_failures; return; }     // per-context error text
	}
	b2r_destroy(c);
}

int main()
{
	std::vector<std::thread> t;
	for (int i = 0; i < 4; ++i) t.emplace_back(worker, i);
	t.emplace_back([] {
		for (int i = 0; i < 200; ++i)
		{
			b2r_context* c = nullptr;
			if (b2r_create(0, 0, 0, &c) == B2R_OK || !b2r_last_error(nullptr)[0]) ++g_failures;      // creation errors are per thread
		}
	});
	for (auto& th : t) th.join();
	printf("threads done, %d failures\n", g_failures.load());
	return g_failures.load() ? 1 : 0;
}
