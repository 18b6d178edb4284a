// TEST INFRASTRUCTURE — force-included (-include) in front of every reference source compiled by
// oracle/Makefile.ref.  It changes no reference file; it only pins two things the reference leaves to
// its environment:
//   * std::random_shuffle (Tracker.cc:483,601) draws from rand(): replaced by the identity permutation,
//     which is what the oracle and the product implement (SURVEY.md §7 hard part 5);
//   * the tracker's usleep() wait for the map-maker THREAD to acknowledge a reset (Tracker.cc:71-74):
//     no thread runs here, so the wait hook performs the pending reset itself (ref_wrap_tracker.cpp).
#pragma once
#include <algorithm>
#include <sstream>
#include <unistd.h>
namespace std { template <class It> inline void ptam_identity_shuffle(It, It) {} }  // the reference spells it std::random_shuffle
#define random_shuffle ptam_identity_shuffle
extern "C" int ptam_ref_usleep(unsigned int usec);
#define usleep ptam_ref_usleep
