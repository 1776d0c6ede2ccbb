/**
 *  @file host_scan.hpp
 *  @brief Host-side scans of caller memory that sit on the critical path of a call (plain C++, compiled by the host
 *         compiler alone so that it may use function multi-versioning).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace szb {

/** Marks the byte values occurring in `bytes[0, count)`: `seen[b] = 1`. Drives the Levenshtein alphabet compaction for
 *  inputs the host scans itself (engine.cu, `resolve_symbols`). */
void mark_present_bytes(uint8_t const *bytes, size_t count, uint32_t (&seen)[256]);

} // namespace szb
