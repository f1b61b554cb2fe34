// stand-in header (test infrastructure, see ../svo_ref_stubs.h)
#pragma once
#include "svo_ref_stubs.h"
