// Shim for the oracle build only. The reference header (Runtime/Tasks/Scheduler.h:1-20) needs the
// MSVC-only <concurrent_queue.h>; the live path-tracer code only needs the containers it re-exports.
#pragma once
#include "Memory/SharedPtr.hpp"
#include "Memory/UniquePtr.hpp"
#include "Containers/Vector.h"
#include "Core/Utils.h"
