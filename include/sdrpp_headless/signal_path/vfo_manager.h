// GUI-LESS TEST HARNESS (include/sdrpp_headless): not part of the drop-in. In the application the reference's own
// signal_path/vfo_manager.{h,cpp} is used unchanged on top of the replaced IQFrontEnd / RxVFO; this header restates its
// surface without ImGui so that tests/cpp/mirror_demo.cpp can drive the path like a module does.
// VFOManager -- host-side mirror of sigpath::vfoManager (reference: core/src/signal_path/vfo_manager.h:6-67,
// vfo_manager.cpp). Named VFO registry pairing a dsp::channel::RxVFO with a (stub) waterfall widget.
#pragma once
#include <cmath>
#include <map>
#include <string>
#include <dsp/channel/rx_vfo.h>
#include "../gui/widgets/waterfall_vfo.h"
#include <utils/event.h>
#include <signal_path/iq_frontend.h>

namespace sigpath { extern IQFrontEnd iqFrontEnd; }

class VFOManager {
public:
    VFOManager() {}

    class VFO {
    public:
        VFO(std::string name, int reference, double offset, double bandwidth, double sampleRate, double minBandwidth,
            double maxBandwidth, bool bandwidthLocked) {
            this->name = name;
            _bandwidth = bandwidth;
            dspVFO = sigpath::iqFrontEnd.addVFO(name, sampleRate, bandwidth, offset);
            wtfVFO = new ImGui::WaterfallVFO;
            wtfVFO->setReference(reference);
            wtfVFO->setBandwidth(bandwidth);
            wtfVFO->setOffset(offset);
            wtfVFO->minBandwidth = minBandwidth;
            wtfVFO->maxBandwidth = maxBandwidth;
            wtfVFO->bandwidthLocked = bandwidthLocked;
            output = dspVFO ? &dspVFO->out : nullptr;
        }
        ~VFO() {
            sigpath::iqFrontEnd.removeVFO(name);
            delete wtfVFO;
        }

        void setOffset(double offset) { wtfVFO->setOffset(offset); dspVFO->setOffset(wtfVFO->centerOffset); }
        double getOffset() { return wtfVFO->generalOffset; }
        void setCenterOffset(double offset) { wtfVFO->setCenterOffset(offset); dspVFO->setOffset(offset); }
        void setBandwidth(double bandwidth, bool updateWaterfall = true) {
            if (_bandwidth == bandwidth) { return; }
            _bandwidth = bandwidth;
            if (updateWaterfall) { wtfVFO->setBandwidth(bandwidth); }
            dspVFO->setBandwidth(bandwidth);
        }
        void setSampleRate(double sampleRate, double bandwidth) {
            dspVFO->setOutSamplerate(sampleRate, bandwidth);
            wtfVFO->setBandwidth(bandwidth);
        }
        void setReference(int ref) { wtfVFO->setReference(ref); }
        void setSnapInterval(double interval) { wtfVFO->setSnapInterval(interval); }
        void setBandwidthLimits(double minBandwidth, double maxBandwidth, bool bandwidthLocked) {
            wtfVFO->minBandwidth = minBandwidth; wtfVFO->maxBandwidth = maxBandwidth; wtfVFO->bandwidthLocked = bandwidthLocked;
        }
        bool getBandwidthChanged(bool erase = true) {
            const bool v = wtfVFO->bandwidthChanged;
            if (erase) { wtfVFO->bandwidthChanged = false; }
            return v;
        }
        double getBandwidth() { return wtfVFO->bandwidth; }
        int getReference() { return wtfVFO->reference; }
        void setColor(ImU32 color) { wtfVFO->color = color; }
        std::string getName() { return name; }

        dsp::stream<dsp::complex_t>* output;
        friend class VFOManager;
        dsp::channel::RxVFO* dspVFO;
        ImGui::WaterfallVFO* wtfVFO;

    private:
        std::string name;
        double _bandwidth;
    };

    VFOManager::VFO* createVFO(std::string name, int reference, double offset, double bandwidth, double sampleRate,
                               double minBandwidth, double maxBandwidth, bool bandwidthLocked) {
        if (name == "" || vfos.find(name) != vfos.end()) { return NULL; }
        auto* vfo = new VFO(name, reference, offset, bandwidth, sampleRate, minBandwidth, maxBandwidth, bandwidthLocked);
        if (!vfo->dspVFO) { delete vfo; return NULL; }
        vfos[name] = vfo;
        onVfoCreated.emit(vfo);
        return vfo;
    }
    void deleteVFO(VFOManager::VFO* vfo) {
        for (auto it = vfos.begin(); it != vfos.end(); ++it) {
            if (it->second != vfo) { continue; }
            const std::string name = it->first;
            onVfoDelete.emit(vfo);
            vfos.erase(it);
            delete vfo;
            onVfoDeleted.emit(name);
            return;
        }
    }

    void setOffset(std::string name, double offset) { if (auto* v = find(name)) { v->setOffset(offset); } }
    double getOffset(std::string name) { auto* v = find(name); return v ? v->getOffset() : 0; }
    void setCenterOffset(std::string name, double offset) { if (auto* v = find(name)) { v->setCenterOffset(offset); } }
    void setBandwidth(std::string name, double bandwidth, bool updateWaterfall = true) { if (auto* v = find(name)) { v->setBandwidth(bandwidth, updateWaterfall); } }
    void setSampleRate(std::string name, double sampleRate, double bandwidth) { if (auto* v = find(name)) { v->setSampleRate(sampleRate, bandwidth); } }
    void setReference(std::string name, int ref) { if (auto* v = find(name)) { v->setReference(ref); } }
    void setBandwidthLimits(std::string name, double minBandwidth, double maxBandwidth, bool bandwidthLocked) {
        if (auto* v = find(name)) { v->setBandwidthLimits(minBandwidth, maxBandwidth, bandwidthLocked); }
    }
    bool getBandwidthChanged(std::string name, bool erase = true) { auto* v = find(name); return v ? v->getBandwidthChanged(erase) : false; }
    double getBandwidth(std::string name) { auto* v = find(name); return v ? v->getBandwidth() : NAN; }
    void setColor(std::string name, ImU32 color) { if (auto* v = find(name)) { v->setColor(color); } }
    int getReference(std::string name) { auto* v = find(name); return v ? v->getReference() : -1; }
    bool vfoExists(std::string name) { return find(name) != nullptr; }

    Event<VFOManager::VFO*> onVfoCreated;
    Event<VFOManager::VFO*> onVfoDelete;
    Event<std::string> onVfoDeleted;

private:
    VFO* find(const std::string& name) { auto it = vfos.find(name); return it == vfos.end() ? nullptr : it->second; }
    std::map<std::string, VFO*> vfos;
};
