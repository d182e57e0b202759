// Minimal Event<T> / EventHandler<T> with the surface the reference's modules use
// (core/src/utils/event.h: bindHandler / unbindHandler / emit; handler = function pointer + ctx).
#pragma once
#include <algorithm>
#include <vector>

template <class T>
struct EventHandler {
    EventHandler() {}
    EventHandler(void (*handler)(T, void*), void* ctx) : handler(handler), ctx(ctx) {}
    void (*handler)(T, void*) = nullptr;
    void* ctx = nullptr;
};

template <class T>
class Event {
public:
    void emit(T value) {
        for (auto* h : handlers) { if (h->handler) { h->handler(value, h->ctx); } }
    }
    void bindHandler(EventHandler<T>* h) { handlers.push_back(h); }
    void unbindHandler(EventHandler<T>* h) { handlers.erase(std::remove(handlers.begin(), handlers.end(), h), handlers.end()); }

private:
    std::vector<EventHandler<T>*> handlers;
};
