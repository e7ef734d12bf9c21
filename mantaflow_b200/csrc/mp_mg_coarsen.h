// Serial coarse-vertex selection for the levels where the parallel closure (k_mg_select) leaves fine vertices with two or
// more undecided interpolation vertices.  Host C++, no CUDA: also compiled on its own by tests/test_mg_coarsen_host.py.
//
// What has to be reproduced (GridMg::genCoarseGrid, multigrid.cpp:520-578) is an ORDER: fine vertices are visited by the
// current number of still-undecided ("free") coarse vertices they interpolate from, fewest first, and among equals the
// vertex whose count changed most recently first.  A visited vertex keeps its first free interpolation vertex (x fastest)
// and drops the others; every decision lowers the count of the <= 27 fine vertices restricting to that coarse vertex.
//
// Here the queue is a set of per-count stacks with lazy deletion: a (vertex, stamp) pair is pushed whenever a count changes,
// and a popped pair is ignored unless its stamp is the vertex's latest one.  The newest valid pair of the lowest non-empty
// stack is exactly "fewest first, most recently changed first".
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace mgcoarsen {

enum : signed char { kInactive = 0, kActive = 1, kFree = 5, kKept = 4, kDropped = 3 };   // = vtInactive, vtActive, vtFree, vtZero, vtRemoved

struct Dim3i { int x, y, z; };

class CountQueue {
public:
	CountQueue(int nVertices, int maxCount) : count_(nVertices, -1), stamp_(nVertices, 0u), stacks_(maxCount + 1), live_(0), clock_(0u) {}
	int live() const { return live_; }
	int countOf(int v) const { return count_[v]; }
	void set(int v, int c) {                      // c < 0 retires the vertex
		if (count_[v] == c) return;
		if (count_[v] < 0) live_++;
		count_[v] = c;
		if (c < 0) { live_--; return; }
		stamp_[v] = ++clock_;
		stacks_[c].push_back(Item{ v, clock_ });
	}
	int takeNext() {                              // precondition: live() > 0
		for (size_t c = 0; c < stacks_.size(); c++) {
			std::vector<Item>& s = stacks_[c];
			while (!s.empty()) {
				const Item it = s.back(); s.pop_back();
				if (count_[it.v] == (int)c && stamp_[it.v] == it.stamp) { count_[it.v] = -1; live_--; return it.v; }
			}
		}
		return -1;
	}
private:
	struct Item { int v; uint32_t stamp; };
	std::vector<int> count_; std::vector<uint32_t> stamp_; std::vector<std::vector<Item>> stacks_;
	int live_; uint32_t clock_;
};

// tf: vertex types of the fine level (only "inactive or not" matters); tc (out): kActive / kInactive per coarse vertex
inline void selectCoarseVertices(Dim3i fine, Dim3i coarse, bool is3D, const std::vector<signed char>& tf, std::vector<signed char>& tc)
{
	const int nf = fine.x * fine.y * fine.z;
	auto fineIdx = [&](int x, int y, int z) { return x + fine.x * (y + fine.y * z); };
	auto coarseIdx = [&](int x, int y, int z) { return x + coarse.x * (y + coarse.y * z); };
	std::fill(tc.begin(), tc.end(), (signed char)kFree);
	CountQueue q(nf, is3D ? 8 : 4);
	for (int v = 0; v < nf; v++) {
		if (tf[v] == kInactive) continue;
		const int x = v % fine.x, y = (v / fine.x) % fine.y, z = v / (fine.x * fine.y);
		q.set(v, 1 << ((x & 1) + (y & 1) + (z & 1)));             // a vertex between coarse vertices interpolates from 2, 4 or 8 of them
	}
	while (q.live() > 0) {
		const int v = q.takeNext();
		const int x = v % fine.x, y = (v / fine.x) % fine.y, z = v / (fine.x * fine.y);
		bool kept = false;
		for (int cz = z / 2; cz <= (z + 1) / 2; cz++) for (int cy = y / 2; cy <= (y + 1) / 2; cy++) for (int cx = x / 2; cx <= (x + 1) / 2; cx++) {
			signed char& t = tc[coarseIdx(cx, cy, cz)];
			if (t != kFree) continue;
			t = kept ? kDropped : kKept;
			kept = true;
			// this coarse vertex is decided: one free interpolation vertex less for every fine vertex around it
			const int x0 = std::max(0, 2 * cx - 1), x1 = std::min(fine.x - 1, 2 * cx + 1);
			const int y0 = std::max(0, 2 * cy - 1), y1 = std::min(fine.y - 1, 2 * cy + 1);
			const int z0 = std::max(0, 2 * cz - 1), z1 = std::min(fine.z - 1, 2 * cz + 1);
			for (int fz = z0; fz <= z1; fz++) for (int fy = y0; fy <= y1; fy++) for (int fx = x0; fx <= x1; fx++) {
				const int f = fineIdx(fx, fy, fz);
				const int c = q.countOf(f);
				if (c >= 0) q.set(f, c > 1 ? c - 1 : -1);
			}
		}
	}
	for (signed char& t : tc) t = (t == kKept) ? kActive : kInactive;
}

}  // namespace mgcoarsen
