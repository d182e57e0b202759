// TEST INFRASTRUCTURE ONLY (oracle). Minimal named-thread wrapper with the constructor shape
// dsp/block.h expects; the oracle never starts block worker threads (it calls process()).
#pragma once
#include <string>
#include <thread>
#include <utility>
namespace threading {
    class thread {
        std::thread _t;
    public:
        thread() = default;
        thread(thread&&) = default;
        thread& operator=(thread&&) = default;
        template <typename F, typename... Args>
        thread(const std::string&, F&& f, Args&&... args) : _t(std::forward<F>(f), std::forward<Args>(args)...) {}
        bool joinable() const { return _t.joinable(); }
        void join() { _t.join(); }
        void detach() { _t.detach(); }
    };
}
