// include/compat/nibtools.h -- `class nib` with the public interface of the reference's src/nibtools.h:12-118 (open / getBase /
// nextBase / size, status codes indexing errormsg[]).  The class lives in breakid_b200/host/nibtools.{h,cc} (it also hands the
// packed payload to the GPU) and is part of libbreakid_compat.so; this header only puts it on the compat include path.
#pragma once
#include "../../breakid_b200/host/nibtools.h"
