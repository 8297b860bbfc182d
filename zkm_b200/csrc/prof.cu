// Device-side timing of kernel families with CUDA events on the launching stream (dev.cuh ProfScope).
// The B200 counterpart of the reference's TimingTree scopes (prover/src/prover.rs:86,144-167,...):
// the reference times wall-clock scopes on the host, here each kernel family is timed on the device.
#include "dev.cuh"
#include <map>
#include <vector>
#include <mutex>
#include <cstdio>

namespace zkm {

namespace {
struct Pending { std::string name; cudaEvent_t e0, e1; double bytes, aux; };
struct Total { double ms = 0; unsigned long long launches = 0; double bytes = 0, aux = 0; };
bool g_on = false;
std::mutex g_mu;                               // worker contexts may profile from several host threads
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_pool;
std::map<std::string, Total> g_totals;
std::vector<std::string> g_order;

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    ZKM_CUDA(cudaEventCreate(&e));
    return e;
}
void resolve() {
    for (Pending& p : g_pending) {
        cudaEventSynchronize(p.e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, p.e0, p.e1);
        auto it = g_totals.find(p.name);
        if (it == g_totals.end()) { g_order.push_back(p.name); it = g_totals.emplace(p.name, Total()).first; }
        it->second.ms += ms; it->second.launches++; it->second.bytes += p.bytes; it->second.aux += p.aux;
        g_pool.push_back(p.e0); g_pool.push_back(p.e1);
    }
    g_pending.clear();
}
}  // namespace

ProfScope::ProfScope(const char* name_, cudaStream_t s_, double b, double aux_) : name(name_), s(s_), bytes(b), aux(aux_) {
    if (!g_on) return;
    std::lock_guard<std::mutex> lk(g_mu);
    e0 = get_event(); e1 = get_event();
    cudaEventRecord(e0, s);
}
ProfScope::~ProfScope() {
    if (!e0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEventRecord(e1, s);
    g_pending.push_back({name, e0, e1, bytes, aux});
    if (g_pending.size() > 8192) resolve();
}
void prof_enable(bool on) { g_on = on; }
void prof_reset() { std::lock_guard<std::mutex> lk(g_mu); resolve(); g_totals.clear(); g_order.clear(); }
bool prof_get(const char* name, double* ms, unsigned long long* launches, double* bytes, double* aux) {
    std::lock_guard<std::mutex> lk(g_mu);
    resolve();
    auto it = g_totals.find(name);
    if (it == g_totals.end()) return false;
    if (ms) *ms = it->second.ms;
    if (launches) *launches = it->second.launches;
    if (bytes) *bytes = it->second.bytes;
    if (aux) *aux = it->second.aux;
    return true;
}
// ---- reference-scope timings (TimedScope, dev.cuh): CUDA events at the scope boundaries on the launching stream, resolved
// when the proof has finished; one record list per host thread (= per context: a context is driven by one thread).
namespace {
bool g_scopes_on = false;
struct ScopeRec { std::string name; int depth; cudaEvent_t e0, e1; };
thread_local std::vector<ScopeRec> t_scopes;
thread_local int t_depth = 0;
thread_local std::string t_last_timing;
}
void scopes_enable(bool on) { g_scopes_on = on; }
TimedScope::TimedScope(const std::string& name, cudaStream_t s_) : s(s_) {
    if (!g_scopes_on) return;
    idx = (int)t_scopes.size();
    ScopeRec r{name, t_depth++, nullptr, nullptr};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        r.e0 = get_event(); r.e1 = get_event();
    }
    cudaEventRecord(r.e0, s);
    t_scopes.push_back(r);
}
TimedScope::~TimedScope() {
    if (idx < 0) return;
    cudaEventRecord(t_scopes[idx].e1, s);
    t_depth--;
}
void scopes_begin() { scopes_finish(); t_last_timing.clear(); }
void scopes_finish() {
    if (t_scopes.empty()) return;
    std::string out;
    for (ScopeRec& r : t_scopes) {
        cudaEventSynchronize(r.e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        char buf[64];
        snprintf(buf, sizeof buf, "%d\t%.4f\t", r.depth, ms);
        out += buf; out += r.name; out += "\n";
        std::lock_guard<std::mutex> lk(g_mu);
        g_pool.push_back(r.e0); g_pool.push_back(r.e1);
    }
    t_scopes.clear();
    t_depth = 0;
    t_last_timing = out;
}
const std::string& scopes_last() { return t_last_timing; }

std::string prof_names() {
    std::lock_guard<std::mutex> lk(g_mu);
    resolve();
    std::string r;
    for (auto& n : g_order) { if (!r.empty()) r += "\n"; r += n; }
    return r;
}

}  // namespace zkm
