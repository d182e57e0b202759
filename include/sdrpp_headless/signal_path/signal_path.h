// GUI-LESS TEST HARNESS (include/sdrpp_headless), see vfo_manager.h beside this file.
// The sigpath:: singletons modules link against (reference: core/src/signal_path/signal_path.h:8-13).
// sourceManager / sinkManager are control plane and out of scope. Define
// SDRPP_SIGPATH_IMPLEMENTATION in exactly one translation unit (the core library).
#pragma once
#include <signal_path/iq_frontend.h>
#include "vfo_manager.h"

namespace sigpath {
    extern IQFrontEnd iqFrontEnd;
    extern VFOManager vfoManager;
}

#ifdef SDRPP_SIGPATH_IMPLEMENTATION
namespace sigpath {
    IQFrontEnd iqFrontEnd;
    VFOManager vfoManager;
}
#endif
