// include/compat/BreakID.h -- lets a translation unit written against the reference's src/BreakID.h (its main(),
// src/BreakID.cc:1-192, first of all) compile against the B200 library instead: same includes a driver needs, the usage
// text, and the stage functions of BreakID_stages.h.  `using namespace std` is kept because the reference's sources rely on it.
#pragma once
#include <getopt.h>

#include <algorithm>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iostream>
#include <list>
#include <sstream>

#include "BreakID_stages.h"
#include "util_bed.h"

using namespace std;

static const string BreakID_help =
    " Usage: \n \t BreakID -i input.bam -o prefix -n nib_folder <options> \n\n "
    "     DESCRIPTION\n "
    "     \t -h -? -help \t help\n "
    "     \t -i*        \t input bam-file\n "
    "     \t -o*        \t output file (prefix only)\n "
    "     \t -n*        \t folder name to nib files\n "
    "     \t -q         \t encompassing reads quality thresholds  [20]\n"
    "     \t -t         \t distance relative to (sqrt(2)*(insert size mean +3* insert size sd))  [2]\n "
    "     \t -fast      \t use the fast cluster strategy [default no] \n "
    "     \t -all       \t no filter enspan out [default is filter]  \n ";
