// include/compat/util_bed.h -- the string helpers of the reference's src/util_bed.h:22-31 that the hot path uses.
// The depth functions of that header (cal_single_base_depth, cal_mean_depth_oc) take samtools file / index handles; here the
// same quantities are computed on the device from the resident records inside findClusterBreakPointInfoSaTag
// (cluster_info::p?_bp_depth; k9_depth in breakid_b200/csrc/bkid_refine.cuh), so they have no host counterpart.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct repeat_str { std::string rawstring; uint16_t start_index; std::string sub_str; uint16_t length; };

// length of the longest run of one character (src/util_bed.cc:224-261); 0 for the empty string (the reference reads out
// of bounds there)
int find_longest_repeat_substring(const std::string &s);
// pieces of s between occurrences of delim, empty pieces dropped; an empty delim returns {s} (src/util_bed.cc:194-222)
std::vector<std::string> split_string(const std::string &s, const std::string &delim);
