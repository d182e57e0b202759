// GUI-LESS TEST HARNESS (include/sdrpp_headless).
// Stub of ImGui::WaterfallVFO: only the state and offset arithmetic that VFOManager::VFO and the
// modules touch (reference: core/src/gui/widgets/waterfall.h, waterfall.cpp:1236-1296). No drawing.
#pragma once
#include <cstdint>
#include <utils/event.h>

typedef uint32_t ImU32;

namespace ImGui {
    class WaterfallVFO {
    public:
        enum { REF_LOWER, REF_CENTER, REF_UPPER, _REF_COUNT };

        // generalOffset is the tuned frequency at the reference edge; centre/lower/upper follow from it
        void setOffset(double offset) {
            generalOffset = offset;
            const double half = bandwidth / 2.0;
            switch (reference) {
            case REF_LOWER: lowerOffset = offset; centerOffset = offset + half; upperOffset = offset + bandwidth; break;
            case REF_UPPER: upperOffset = offset; centerOffset = offset - half; lowerOffset = offset - bandwidth; break;
            default: centerOffset = offset; lowerOffset = offset - half; upperOffset = offset + half; break;
            }
            centerOffsetChanged = lowerOffsetChanged = upperOffsetChanged = true;
        }
        void setCenterOffset(double offset) {
            const double half = bandwidth / 2.0;
            generalOffset = (reference == REF_LOWER) ? offset - half : (reference == REF_UPPER) ? offset + half : offset;
            centerOffset = offset; lowerOffset = offset - half; upperOffset = offset + half;
            centerOffsetChanged = lowerOffsetChanged = upperOffsetChanged = true;
        }
        void setBandwidth(double bw) {
            if (bw == bandwidth || bw < 0) { return; }
            bandwidth = bw;
            const double half = bw / 2.0;
            switch (reference) {
            case REF_LOWER: centerOffset = lowerOffset + half; upperOffset = lowerOffset + bw; centerOffsetChanged = true; break;
            case REF_UPPER: centerOffset = upperOffset - half; lowerOffset = upperOffset - bw; centerOffsetChanged = true; break;
            default: lowerOffset = centerOffset - half; upperOffset = centerOffset + half; break;
            }
            bandwidthChanged = true;
        }
        void setReference(int ref) {
            if (ref == reference || ref < 0 || ref >= _REF_COUNT) { return; }
            reference = ref;
            setOffset(generalOffset);
        }
        void setSnapInterval(double interval) { snapInterval = interval; }

        double generalOffset = 0, centerOffset = 0, lowerOffset = 0, upperOffset = 0;
        double bandwidth = 1, snapInterval = 5000; // defaults of the reference widget (waterfall.h:38-39)
        int reference = REF_CENTER;
        double minBandwidth = 0, maxBandwidth = 0;
        bool bandwidthLocked = false;
        bool centerOffsetChanged = false, lowerOffsetChanged = false, upperOffsetChanged = false, bandwidthChanged = false;
        ImU32 color = 0xFFFFFFFFu;
        Event<double> onUserChangedBandwidth;
    };
}
