// dsp::block: start/stop/tempStop control surface of a processing block
// (contract of the reference's core/src/dsp/block.h:19-133, re-implemented on std::thread).
#pragma once
#include <algorithm>
#include <cassert>
#include <mutex>
#include <thread>
#include <vector>
#include "stream.h"
#include "types.h"

namespace dsp {
    class generic_block {
    public:
        virtual ~generic_block() {}
        virtual void start() {}
        virtual void stop() {}
        virtual int run() { return -1; }
    };

    class block : public generic_block {
    public:
        virtual ~block() {
            if (!_block_init) { return; }
            stop();
            _block_init = false;
        }

        // idempotent; serialised by ctrlMtx like every setter of a derived block
        virtual void start() {
            assert(_block_init);
            std::lock_guard<std::recursive_mutex> lck(ctrlMtx);
            if (running) { return; }
            running = true;
            doStart();
        }
        virtual void stop() {
            assert(_block_init);
            std::lock_guard<std::recursive_mutex> lck(ctrlMtx);
            if (!running) { return; }
            doStop();
            running = false;
        }

        // nestable pause used by setters: the worker is joined on the first tempStop and respawned by the
        // matching tempStart
        void tempStop() {
            assert(_block_init);
            std::lock_guard<std::recursive_mutex> lck(ctrlMtx);
            if (tempStopDepth++ > 0) { return; }
            if (running && !tempStopped) { doStop(); tempStopped = true; }
        }
        void tempStart() {
            assert(_block_init);
            std::lock_guard<std::recursive_mutex> lck(ctrlMtx);
            if (tempStopDepth == 0 || --tempStopDepth > 0) { return; }
            if (tempStopped) { doStart(); tempStopped = false; }
        }

        virtual int run() = 0; // < 0 ends the worker

    protected:
        void workerLoop() { while (run() >= 0) {} }

        virtual void doStart() { workerThread = std::thread(&block::workerLoop, this); }
        virtual void doStop() {
            for (auto* s : inputs) { s->stopReader(); }
            for (auto* s : outputs) { s->stopWriter(); }
            if (workerThread.joinable()) { workerThread.join(); }
            for (auto* s : inputs) { s->clearReadStop(); }
            for (auto* s : outputs) { s->clearWriteStop(); }
        }

        void acquire() { ctrlMtx.lock(); }
        void release() { ctrlMtx.unlock(); }
        void registerInput(untyped_stream* s) { inputs.push_back(s); }
        void unregisterInput(untyped_stream* s) { inputs.erase(std::remove(inputs.begin(), inputs.end(), s), inputs.end()); }
        void registerOutput(untyped_stream* s) { outputs.push_back(s); }
        void unregisterOutput(untyped_stream* s) { outputs.erase(std::remove(outputs.begin(), outputs.end(), s), outputs.end()); }

        bool _block_init = false;
        std::recursive_mutex ctrlMtx;
        std::vector<untyped_stream*> inputs, outputs;
        bool running = false, tempStopped = false;
        int tempStopDepth = 0;
        std::thread workerThread;
    };
}
