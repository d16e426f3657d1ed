// CPU check of hb_order.h (PassOrder): worker threads take batches from a counter and finish them after random delays; what they commit must come out
// in batch order, and a failing batch must release every waiter.    g++ -O2 -std=c++17 -pthread tests/cpp/order_test.cpp -o order_test
#include "../../hifiasm_b200/csrc/hb_order.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>

static int run(size_t n, int n_threads, long fail_at, std::vector<size_t> &seq, std::vector<int> &rcs)
{
	PassOrder order(n); std::atomic<size_t> nextk(0); seq.clear(); rcs.assign(n_threads, 0);
	auto worker = [&](int t) {
		std::mt19937 rng(1234 + t);
		for (;;) {
			const size_t k = nextk.fetch_add(1); if (k >= n) break;
			{ std::lock_guard<std::mutex> lk(order.mu); if (order.aborted) break; }
			std::this_thread::sleep_for(std::chrono::microseconds(rng() % 300)); // the batch's "work": later batches often finish first
			const int rc = order.commit(k, [&]() -> int { if ((long)k == fail_at) return -5; seq.push_back(k); return 0; });
			if (rc) { rcs[t] = rc; order.abort(); break; }
		}
	};
	std::vector<std::thread> th; for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t);
	for (auto &t : th) t.join();
	return 0;
}

int main()
{
	std::vector<size_t> seq; std::vector<int> rcs;
	for (int threads = 1; threads <= 4; threads++) {
		run(200, threads, -1, seq, rcs);
		if (seq.size() != 200) { printf("FAIL: %zu commits with %d threads\n", seq.size(), threads); return 1; }
		for (size_t i = 0; i < seq.size(); i++) if (seq[i] != i) { printf("FAIL: commit %zu is batch %zu (%d threads)\n", i, seq[i], threads); return 1; }
		for (int rc : rcs) if (rc) { printf("FAIL: rc %d without a failure\n", rc); return 1; }
	}
	for (int threads = 2; threads <= 4; threads++) {
		run(200, threads, 57, seq, rcs);
		if (seq.size() != 57) { printf("FAIL: %zu commits before the failing batch 57 (%d threads)\n", seq.size(), threads); return 1; }
		for (size_t i = 0; i < seq.size(); i++) if (seq[i] != i) { printf("FAIL: order before the failure\n"); return 1; }
		int own = 0, released = 0; for (int rc : rcs) { if (rc == -5) own++; else if (rc == HB_E_ABORTED) released++; else if (rc) { printf("FAIL: rc %d\n", rc); return 1; } }
		if (own != 1) { printf("FAIL: %d threads report the failure itself\n", own); return 1; }
		(void)released; // (threads that were waiting get HB_E_ABORTED; threads between batches just stop)
	}
	printf("OK\n");
	return 0;
}
