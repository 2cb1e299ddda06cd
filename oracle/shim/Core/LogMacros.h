// Shim for the oracle build only (test infrastructure, never linked into the product).
// Replaces /root/reference/Runtime/Core/LogMacros.h:1-60, which pulls in <windows.h> and the Editor submodule.
#pragma once
#include <cstdio>
#define SAILOR_LOG(...)       do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define SAILOR_LOG_ERROR(...) SAILOR_LOG(__VA_ARGS__)
