// dsp::stream<T>: the blocking double-buffer hand-off between one writer and one reader
// (contract of the reference's core/src/dsp/stream.h:24-140, re-implemented). Replaces that file in the source overlay
// (tools/make_overlay.py).
//
// Differences that matter for the GPU path: both buffers are PINNED host memory
// (sdrpp_cuda_host_alloc) so an IQ block can be DMA'd to the device straight out of readBuf; if no
// CUDA device is present the allocation falls back to ordinary aligned memory so that host-only code
// (and CPU tests) still work -- compute calls still fail loudly without a GPU.
#pragma once
#include <string.h>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <volk/volk.h>      // like the reference's stream.h: the rest of dsp/ relies on these two arriving through it
#include "buffer/buffer.h"
#include <sdrpp_cuda.h>

#define STREAM_BUFFER_SIZE 1000000 // elements per buffer, as in the reference (stream.h:9)

namespace dsp {
    class untyped_stream {
    public:
        virtual ~untyped_stream() {}
        virtual bool swap(int size) { (void)size; return false; }
        virtual int read() { return -1; }
        virtual void flush() {}
        virtual void stopWriter() {}
        virtual void clearWriteStop() {}
        virtual void stopReader() {}
        virtual void clearReadStop() {}
    };

    namespace detail {
        struct HostBuf {
            void* p = nullptr;
            bool pinned = false;
            void alloc(size_t bytes) {
                release();
                p = sdrpp_cuda_device_count() > 0 ? sdrpp_cuda_host_alloc(bytes) : nullptr;
                pinned = (p != nullptr);
                if (!p) { p = std::aligned_alloc(64, (bytes + 63) & ~(size_t)63); }
            }
            void release() {
                if (!p) { return; }
                if (pinned) { sdrpp_cuda_host_free(p); } else { std::free(p); }
                p = nullptr;
            }
        };
    }

    template <class T>
    class stream : public untyped_stream {
    public:
        stream() { setBufferSize(STREAM_BUFFER_SIZE); }
        virtual ~stream() { free(); }

        virtual void setBufferSize(int samples) {
            bufs[0].alloc(sizeof(T) * (size_t)samples);
            bufs[1].alloc(sizeof(T) * (size_t)samples);
            writeBuf = (T*)bufs[0].p;
            readBuf = (T*)bufs[1].p;
        }

        // Writer: hand the filled writeBuf (size elements) to the reader. Blocks until the reader has flushed
        // the previous block; returns false if the writer was stopped.
        virtual bool swap(int size) {
            std::unique_lock<std::mutex> lck(mtx);
            cv.wait(lck, [this] { return state == EMPTY || writerStop; });
            if (writerStop) { return false; }
            T* t = writeBuf; writeBuf = readBuf; readBuf = t;
            dataSize = size;
            state = FULL;
            lck.unlock();
            cv.notify_all();
            return true;
        }

        // Reader: wait for a block; returns its element count or -1 if the reader was stopped.
        virtual int read() {
            std::unique_lock<std::mutex> lck(mtx);
            cv.wait(lck, [this] { return state == FULL || readerStop; });
            return readerStop ? -1 : dataSize;
        }

        // Reader: done with readBuf.
        virtual void flush() {
            { std::lock_guard<std::mutex> lck(mtx); state = EMPTY; }
            cv.notify_all();
        }

        virtual void stopWriter() { { std::lock_guard<std::mutex> lck(mtx); writerStop = true; } cv.notify_all(); }
        virtual void clearWriteStop() { std::lock_guard<std::mutex> lck(mtx); writerStop = false; }
        virtual void stopReader() { { std::lock_guard<std::mutex> lck(mtx); readerStop = true; } cv.notify_all(); }
        virtual void clearReadStop() { std::lock_guard<std::mutex> lck(mtx); readerStop = false; }

        void free() {
            bufs[0].release(); bufs[1].release();
            writeBuf = nullptr; readBuf = nullptr;
        }

        T* writeBuf = nullptr;
        T* readBuf = nullptr;

    private:
        enum State { EMPTY, FULL };
        detail::HostBuf bufs[2];
        std::mutex mtx;
        std::condition_variable cv;
        State state = EMPTY;
        bool writerStop = false, readerStop = false;
        int dataSize = 0;
    };
}
