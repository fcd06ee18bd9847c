// Reverse-mode AD over the device ops: a slim host tape replacing stan::math's arena + reverse_pass_callback for the Tensor rules of the reference
// (include/Xped/AD/ADTensor.hpp:137-150 adjoint, :171-185 permute, :193-219 contract, :237-374 tSVD + truncation, :492-602 norm / maxNorm / trace /
// diag_sqrt / diag_inv / twist, :640-763 + / scalar * / operator*) and the checkpoint replay of the CTM driver (src/PEPS/CTM.cpp:399-419,976-987,
// 1327-1335,1407-1416,1751-1758: a region is run without recording and re-run under a sub-tape during the reverse sweep).
//
// Values are identified by their device buffer (Tensor::id()); the tape keeps aliases of the values its reverse rules need, so the buffers live until
// the tape is destroyed.  Adjoints are plain element-wise derivatives of the loss w.r.t. the stored `m_data` image (no degeneracy weights), like the
// reference's `adj()` tensors.  The loss is a weighted sum of traces: seed_trace(v, w) adds w * d trace(v) / d v = w * WeightedIdentity.
#pragma once
#include <functional>
#include <memory>
#include <unordered_map>
#include <vector>

#include "engine.hpp"

namespace xb {

using TP = std::shared_ptr<Tensor>;

class Tape
{
public:
    explicit Tape(Context& c) : ctx_(c) {}
    bool recording = true;

    void record(std::function<void(Tape&)> back) { nodes_.push_back(std::move(back)); }
    Tensor* keep(const Tensor& v);                   // alias owned by the tape (valid until the tape dies)
    Tensor* grad(const Tensor& v) const;             // accumulated adjoint or nullptr
    void accumulate(const Tensor& v, Tensor* g);     // adj(v) += g, takes ownership of g
    void seed_trace(const Tensor& v, double w);      // adj(v) += w * WeightedIdentity  (ADTensor.hpp:537-542)
    void seed(const Tensor& v, const Tensor& g);     // adj(v) += g (copied)
    void backward();                                 // runs the recorded rules in reverse order (recording is off meanwhile); nodes are consumed
    size_t size() const { return nodes_.size(); }

    // Checkpointed region: fn(inputs) is evaluated WITHOUT recording; during the reverse sweep it is evaluated again under a sub-tape whose output
    // adjoints are seeded from this tape, and the input adjoints flow back here.  Trades one extra forward evaluation for not storing the region's intermediates.
    std::vector<TP> checkpoint(const std::vector<TP>& inputs, std::function<std::vector<TP>(const std::vector<TP>&)> fn);

private:
    Context& ctx_;
    std::vector<std::function<void(Tape&)>> nodes_;
    std::unordered_map<const DevBuf*, std::unique_ptr<Tensor>> grads_;
    std::vector<std::unique_ptr<Tensor>> keep_;
};

// RAII: no recording inside compound ops whose rule is recorded as one node
struct TapePause
{
    Tape* t;
    bool saved = false;
    explicit TapePause(Context& c) : t(c.tape)
    {
        if(t) {
            saved = t->recording;
            t->recording = false;
        }
    }
    ~TapePause()
    {
        if(t) t->recording = saved;
    }
};

}   // namespace xb
