// TEST INFRASTRUCTURE ONLY (oracle). Silent stand-in for the reference's logger so its dsp/
// headers compile without the application runtime (core/src/utils/flog.h is not on the path).
#pragma once
#include <exception>
namespace flog {
    template <class... A> inline void debug(const char*, A...) {}
    template <class... A> inline void info(const char*, A...) {}
    template <class... A> inline void warn(const char*, A...) {}
    template <class... A> inline void error(const char*, A...) {}
    inline void exception(const std::exception&) {}
    inline void exception() {}
}
