// teigh / tQR / SVD reverse rule / MPS transfer contractions. (Filled in after the first GPU bring-up.)
#include <stdexcept>

#include "engine.hpp"

namespace xb {

void Context::tqr(const Tensor&, Tensor**, Tensor**) { throw std::logic_error("tQR: not implemented yet"); }
void Context::teigh(const Tensor&, Tensor**, Tensor**) { throw std::logic_error("teigh: not implemented yet"); }
Tensor* Context::tsvd_backward(const Tensor&, const Tensor&, const Tensor&, const Tensor&, const Tensor*, const Tensor*, const Tensor*)
{
    throw std::logic_error("tsvd_backward: not implemented yet");
}
void Context::contract_L(const Tensor&, const Tensor&, const Tensor&, Tensor**) { throw std::logic_error("contract_L: not implemented yet"); }
void Context::contract_R(const Tensor&, const Tensor&, const Tensor&, Tensor**) { throw std::logic_error("contract_R: not implemented yet"); }

}   // namespace xb
