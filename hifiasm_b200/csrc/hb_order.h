// hb_order.h — results of concurrently processed batches, committed in batch order (plain C++: also compiled by the CPU tests).
//
// The batches of a pass run on several lanes (engine.cu: run_batches); what a batch adds to the pass's host-side results — lists, edit scripts, running
// offsets — must not depend on which lane finished first: commit(k, fn) runs fn when the batches before k have committed.  A failing batch (its fn
// returns non-zero, or abort()) releases everybody who waits: they return HB_E_ABORTED and the failing batch's own error is the one reported.
#pragma once
#include <mutex>
#include <condition_variable>
#include <vector>
#include <cstddef>

#define HB_E_ABORTED (-101) /* internal: the batch gave up because another one failed */
struct PassOrder {
	std::mutex mu; std::condition_variable cv; size_t next; bool aborted; std::vector<char> done;
	explicit PassOrder(size_t n) : next(0), aborted(false), done(n, 0) {}
	template <typename F> int commit(size_t k, F fn)
	{
		std::unique_lock<std::mutex> lk(mu);
		cv.wait(lk, [&]() { return next == k || aborted; });
		if (aborted) return HB_E_ABORTED; // (another batch failed: its error is the one reported)
		const int rc = fn();
		if (rc) aborted = true; else { done[k] = 1; next = k + 1; }
		lk.unlock(); cv.notify_all();
		return rc;
	}
	void abort() { { std::lock_guard<std::mutex> lk(mu); aborted = true; } cv.notify_all(); }
};
