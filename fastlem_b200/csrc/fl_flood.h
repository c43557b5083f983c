// fl_flood.h -- host-side flood order of lake removal (see fl_flood.cpp).
#pragma once
#include <cstdint>

#define FL_RANK_NONE 0xFFFFFFFFu

// rank[i] = sequence number of node i's first pop in the flood of reference
// src/lem/stream_tree.rs:175-243 (FL_RANK_NONE if the flood never reaches i).
void fl_flood_rank(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                   const uint32_t* outlets, uint32_t n_outlets, uint32_t* rank);

// The same replay, stopped after `stop_after` nodes have been ranked; returns how many were.  With positive edge
// lengths every outlet (key 0.0) is popped before any other node, so stop_after = number of distinct outlets
// yields exactly the outlets' ranks -- the part of the order that depends on the heap's tie behaviour.
uint32_t fl_flood_rank_prefix(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                              const uint32_t* outlets, uint32_t n_outlets, uint32_t* rank, uint32_t stop_after);
