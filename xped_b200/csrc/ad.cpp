// Tape mechanics (see ad.hpp); the reverse rules themselves are recorded next to the forward ops (ops.cpp, linalg.cu, ctm.cpp).
#include "ad.hpp"

#include <stdexcept>

namespace xb {

bool Context::taping() const { return tape != nullptr && tape->recording; }

Tensor* Tape::keep(const Tensor& v)
{
    keep_.emplace_back(v.alias());
    return keep_.back().get();
}

Tensor* Tape::grad(const Tensor& v) const
{
    auto it = grads_.find(v.id());
    return it == grads_.end() ? nullptr : it->second.get();
}

void Tape::accumulate(const Tensor& v, Tensor* g)
{
    std::unique_ptr<Tensor> gp(g);
    if(gp->layout->size != v.layout->size || gp->layout->sector != v.layout->sector) {
        // an adjoint that came out of a product can lack sectors the value has (or list them in another order): move the blocks by sector label
        const Layout &Lg = *gp->layout, &Lv = *v.layout;
        std::unique_ptr<Tensor> fit(new Tensor(&ctx_, v.layout, true));
        for(size_t i = 0; i < Lv.sector.size(); ++i) {
            const int j = Lg.find(Lv.sector[i]);
            if(j < 0) continue;
            if(Lg.rows[j] != Lv.rows[i] || Lg.cols[j] != Lv.cols[i]) throw std::logic_error("AD: adjoint block has a different shape than its value's");
            if(Lv.rows[i] * Lv.cols[i] == 0) continue;
            check_cuda(cudaMemcpyAsync(fit->d + Lv.off[i], gp->d + Lg.off[j], sizeof(double) * Lv.rows[i] * Lv.cols[i], cudaMemcpyDeviceToDevice, ctx_.stream),
                       "AD adjoint relayout");
        }
        gp = std::move(fit);
    }
    auto it = grads_.find(v.id());
    if(it == grads_.end()) {
        keep(v);   // the key must stay a live buffer
        grads_.emplace(v.id(), std::move(gp));
        return;
    }
    const bool rec = recording;
    recording = false;
    std::unique_ptr<Tensor> sum(ctx_.axpby(1.0, *it->second, 1.0, gp.get()));
    recording = rec;
    it->second = std::move(sum);
}

void Tape::seed_trace(const Tensor& v, double w)
{
    const bool rec = recording;
    recording = false;
    std::unique_ptr<Tensor> id(ctx_.weighted_identity(v.layout));
    Tensor* g = ctx_.axpby(w, *id, 0.0, nullptr);
    recording = rec;
    accumulate(v, g);
}

void Tape::seed(const Tensor& v, const Tensor& g)
{
    const bool rec = recording;
    recording = false;
    Tensor* c = ctx_.copy(g);
    recording = rec;
    accumulate(v, c);
}

void Tape::backward()
{
    const bool rec = recording;
    recording = false;
    Tape* outer = ctx_.tape;
    ctx_.tape = this;
    try {
        while(!nodes_.empty()) {
            auto fn = std::move(nodes_.back());
            nodes_.pop_back();
            fn(*this);
        }
    } catch(...) {
        ctx_.tape = outer;
        recording = rec;
        throw;
    }
    ctx_.tape = outer;
    recording = rec;
}

std::vector<TP> Tape::checkpoint(const std::vector<TP>& inputs, std::function<std::vector<TP>(const std::vector<TP>&)> fn)
{
    const bool rec = recording;
    recording = false;
    std::vector<TP> outs;
    try {
        outs = fn(inputs);
    } catch(...) {
        recording = rec;
        throw;
    }
    recording = rec;
    if(!rec) return outs;
    std::vector<Tensor*> in_keep, out_keep;
    for(const TP& t : inputs) in_keep.push_back(keep(*t));
    for(const TP& t : outs) out_keep.push_back(keep(*t));
    Context* ctx = &ctx_;
    record([ctx, in_keep, out_keep, fn](Tape& outer) {
        // an output that IS one of the inputs (passed through unchanged) already carries its adjoint under the input's identity: nothing to replay for it
        std::vector<bool> pass(out_keep.size(), false);
        bool any = false;
        for(size_t i = 0; i < out_keep.size(); ++i) {
            for(Tensor* t : in_keep) pass[i] = pass[i] || t->id() == out_keep[i]->id();
            any = any || (!pass[i] && outer.grad(*out_keep[i]) != nullptr);
        }
        if(!any) return;
        // replay under a sub-tape
        std::vector<TP> ins;
        for(Tensor* t : in_keep) ins.emplace_back(t->alias());
        Tape sub(*ctx);
        Tape* saved = ctx->tape;
        ctx->tape = &sub;
        std::vector<TP> outs2;
        try {
            outs2 = fn(ins);
        } catch(...) {
            ctx->tape = saved;
            throw;
        }
        ctx->tape = saved;
        if(outs2.size() != out_keep.size()) throw std::logic_error("AD checkpoint: the replay returned a different number of outputs");
        for(size_t i = 0; i < outs2.size(); ++i) {
            if(pass[i]) continue;
            if(Tensor* g = outer.grad(*out_keep[i])) sub.seed(*outs2[i], *g);
        }
        sub.backward();
        for(size_t i = 0; i < ins.size(); ++i)
            if(Tensor* g = sub.grad(*ins[i])) outer.accumulate(*in_keep[i], ctx->copy(*g));
    });
    return outs;
}

}   // namespace xb
