// SmpcController.cuh (shim) -- the reference's header name (/root/reference/src/SmpcController.cuh) for callers compiled against rapidnet-b200:
// class SmpcController of rapidnet_b200/host/rapidnet_host.hpp in the global namespace, where the reference declares it.
#pragma once
#include "Configuration.h"
#include "Engine.cuh"
using rapidnet::SmpcController;
