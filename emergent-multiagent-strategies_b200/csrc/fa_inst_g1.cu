#include "fa_inst.cuh"
namespace fa {
FA_INSTANTIATE(1)
}
