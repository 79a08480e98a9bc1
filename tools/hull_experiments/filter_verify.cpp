// CPU experiment: "filter + verify" for the monotone chains of convex_hull::andrew with its tolerance pop test.
// Points far above a cheap upper bound of the lower hull (polyline through per-block minima) are left out of the sequential
// machine; afterwards every left-out run is replayed from the recorded stack top (a, b) to prove it was a net no-op:
// no point of the run ever pops b, and the next kept point pops every run member that is still on the stack.
// Reports the kept fraction, how many chains verify, and checks that verified chains equal the full machine's output.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
struct P { float x, y; };
#define M 1e-4f
struct L { float l0, l1, l2; };
static L line(P a, P b) { return {a.y * b.x - a.x * b.y, b.y - a.y, a.x - b.x}; }
static float side(const L& l, P p) { return (l.l0 + p.x * l.l1) + p.y * l.l2; }
static bool pops(P a, P b, P p) { return side(line(a, b), p) <= M; }
static void full(const std::vector<P>& pts, std::vector<P>& st) {
    for (P p : pts) { while (st.size() > 1 && pops(st[st.size() - 2], st[st.size() - 1], p)) st.pop_back(); st.push_back(p); }
}
int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 64;
    const float delta_rel = argc > 2 ? atof(argv[2]) : 0.05f;
    FILE* f = fopen("/tmp/hx/pts.bin", "rb");
    uint32_t ns; fread(&ns, 4, 1, f);
    long chains = 0, verified = 0, kept_total = 0, pts_total = 0, equal = 0, runs_total = 0, maxrun = 0, fail_b = 0, fail_k = 0;
    for (uint32_t s = 0; s < ns; ++s) {
        uint32_t n; fread(&n, 4, 1, f);
        std::vector<P> pts(n); fread(pts.data(), 8, n, f);
        for (int dir = 0; dir < 2; ++dir) {
            if (dir) std::reverse(pts.begin(), pts.end());
            // In the machine's frame "up" is the side of positive turn. For the ascending (lower) chain positive turn = above; for the descending chain = below.
            const float sgn = dir ? -1.f : 1.f;
            std::vector<P> truth; full(pts, truth);
            // envelope: per-block extreme (lowest in machine frame) points
            const size_t bs = (n + K - 1) / K;
            std::vector<P> env; env.push_back(pts[0]);
            float ymin = 1e30f, ymax = -1e30f;
            for (P p : pts) { ymin = std::min(ymin, p.y); ymax = std::max(ymax, p.y); }
            const float delta = delta_rel * (ymax - ymin);
            for (size_t b0 = 0; b0 < n; b0 += bs) {
                size_t b1 = std::min<size_t>(n, b0 + bs); size_t best = b0;
                for (size_t i = b0; i < b1; ++i) if (sgn * pts[i].y < sgn * pts[best].y) best = i;
                if (best != 0 && best != n - 1) env.push_back(pts[best]);
            }
            env.push_back(pts[n - 1]);
            // keep flags
            std::vector<char> keep(n, 0);
            size_t e = 0;
            for (size_t i = 0; i < n; ++i) {
                const P p = pts[i];
                // advance the envelope segment: env is ordered like pts (ascending or descending x)
                while (e + 2 < env.size() && (dir ? env[e + 1].x >= p.x : env[e + 1].x <= p.x)) ++e;
                const P u = env[e], v = env[e + 1];
                float ye;
                if (u.x == v.x) ye = sgn * std::min(sgn * u.y, sgn * v.y);
                else { float t = (p.x - u.x) / (v.x - u.x); t = std::min(1.f, std::max(0.f, t)); ye = u.y + t * (v.y - u.y); }
                keep[i] = sgn * p.y <= sgn * ye + delta;
            }
            keep[0] = keep[1] = keep[n - 1] = 1;
            // filtered machine with the top two recorded after every kept point
            std::vector<P> st; std::vector<size_t> kidx; std::vector<P> ra, rb;
            for (size_t i = 0; i < n; ++i) if (keep[i]) {
                P p = pts[i];
                while (st.size() > 1 && pops(st[st.size() - 2], st[st.size() - 1], p)) st.pop_back();
                st.push_back(p);
                kidx.push_back(i);
                ra.push_back(st.size() > 1 ? st[st.size() - 2] : p); rb.push_back(p);
            }
            // verification of the runs between kept points
            bool ok = true;
            for (size_t j = 0; j + 1 < kidx.size() && ok; ++j) {
                const size_t i0 = kidx[j] + 1, i1 = kidx[j + 1];
                if (i0 == i1) continue;
                runs_total++; maxrun = std::max<long>(maxrun, i1 - i0);
                if (j == 0) { ok = false; break; }   // (cannot happen: points 0 and 1 are kept)
                const P a = ra[j], b = rb[j];
                std::vector<P> us;   // run members on the stack, above b
                for (size_t i = i0; i < i1 && ok; ++i) {
                    const P p = pts[i];
                    while (!us.empty()) { const P lo = us.size() > 1 ? us[us.size() - 2] : b; if (!pops(lo, us.back(), p)) break; us.pop_back(); }
                    if (us.empty() && pops(a, b, p)) { ok = false; fail_b++; }
                    us.push_back(p);
                }
                const P k = pts[i1];
                while (ok && !us.empty()) { const P lo = us.size() > 1 ? us[us.size() - 2] : b; if (!pops(lo, us.back(), k)) { ok = false; fail_k++; } else us.pop_back(); }
            }
            chains++; pts_total += n; kept_total += kidx.size();
            if (ok) { verified++; bool eq = st.size() == truth.size(); for (size_t i = 0; eq && i < st.size(); ++i) eq = st[i].x == truth[i].x && st[i].y == truth[i].y; equal += eq; }
        }
    }
    printf("K=%d delta=%.3f: kept %.1f%% of points; %ld of %ld chains verify (%ld equal to the full machine); runs %ld (max %ld); fails: pops-b %ld, kept-does-not-pop %ld\n",
           K, delta_rel, 100.0 * kept_total / pts_total, verified, chains, equal, runs_total, maxrun, fail_b, fail_k);
}
