#pragma once
#include "../../include/bya.h"
namespace bya {
using GemmArgs = ::ByaGemmArgs;
}
