// TEST INFRASTRUCTURE ONLY -- see sift_oracle.h.  Parity status: PINNED (bit-exact against
// oracle/_ref/libsift_ref.so, the real reference, by tests/test_oracle.py; golden vectors in
// tests/golden/ come from that reference, not from this file).
//
// Plain FP64 restatement of the reference SIFT.  Floating-point expressions keep the
// reference's association order so that results are bit-identical on the same libm; build with
// -ffp-contract=off semantics (x86-64 baseline has no FMA, matching the reference's -O3 build).

#include "sift_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

constexpr double kPi = 3.14159265358979323846;  // M_PI
constexpr double kTwoPi = 6.283185307179586;    // M_PI2, sift.hh:5
constexpr int kMaxSteps = 5;                    // sift.hh:7

struct Plane {
    int w = 0, h = 0;
    std::vector<double> v;
    Plane() {}
    Plane(int w_, int h_) : w(w_), h(h_), v((size_t)w_ * h_) {}
    double at(int x, int y) const { return v[(size_t)y * w + x]; }
    double& at(int x, int y) { return v[(size_t)y * w + x]; }
};

// image.cpp:8-24 -- BT.709 luma, summed left to right; single-channel input passes through.
Plane to_gray(const double* px, int w, int h, int c) {
    Plane g(w, h);
    if (c == 1) {
        std::memcpy(g.v.data(), px, sizeof(double) * (size_t)w * h);
        return g;
    }
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const double* p = px + i * c;
        g.v[i] = 0.2126 * p[0] + 0.7152 * p[1] + 0.0722 * p[2];
    }
    return g;
}

// image.cpp:62-88 with fx = fy = 2: source coordinate i/2, right/bottom neighbour clamped.
Plane upsample2(const Plane& s) {
    Plane d(s.w * 2, s.h * 2);
    for (int j = 0; j < d.h; ++j) {
        double fy = j / 2.0;
        int y0 = (int)fy, y1 = std::min(y0 + 1, s.h - 1);
        double ty = fy - y0;
        for (int i = 0; i < d.w; ++i) {
            double fx = i / 2.0;
            int x0 = (int)fx, x1 = std::min(x0 + 1, s.w - 1);
            double tx = fx - x0;
            double top = s.at(x0, y0) * (1 - tx) + s.at(x1, y0) * tx;
            double bot = s.at(x0, y1) * (1 - tx) + s.at(x1, y1) * tx;
            d.at(i, j) = top * (1 - ty) + bot * ty;
        }
    }
    return d;
}

// image.cpp:226-235 -- half kernel, ceil(3 sigma)+1 taps, un-normalised Gaussian values.
std::vector<double> gaussian_taps(double sigma) {
    int n = (int)std::ceil(3 * sigma) + 1;
    std::vector<double> k(n);
    double denom = 2 * sigma * sigma;
    double coef = 1 / (std::sqrt(2 * kPi) * sigma);
    for (int i = 0; i < n; ++i) k[i] = std::exp(-i * i / denom) * coef;
    return k;
}

// image.cpp:156-214 -- horizontal then vertical pass; clamp-to-edge; symmetric pair sums;
// each output divided by the tap-weight total accumulated alongside.
Plane blur(const Plane& src, double sigma) {
    std::vector<double> k = gaussian_taps(sigma);
    const int n = (int)k.size();
    Plane tmp(src.w, src.h), dst(src.w, src.h);
    for (int y = 0; y < src.h; ++y)
        for (int x = 0; x < src.w; ++x) {
            double acc = src.at(x, y) * k[0], wsum = k[0];
            for (int u = 1; u < n; ++u) {
                int xr = std::min(x + u, src.w - 1), xl = std::max(x - u, 0);
                acc += k[u] * (src.at(xr, y) + src.at(xl, y));
                wsum += 2.0 * k[u];
            }
            tmp.at(x, y) = acc / wsum;
        }
    for (int y = 0; y < src.h; ++y)
        for (int x = 0; x < src.w; ++x) {
            double acc = tmp.at(x, y) * k[0], wsum = k[0];
            for (int u = 1; u < n; ++u) {
                int yd = std::min(y + u, src.h - 1), yu = std::max(y - u, 0);
                acc += k[u] * (tmp.at(x, yd) + tmp.at(x, yu));
                wsum += 2.0 * k[u];
            }
            dst.at(x, y) = acc / wsum;
        }
    return dst;
}

// image.cpp:41-55 -- keep every second pixel of every second row, floor dims.
Plane decimate(const Plane& s) {
    Plane d(s.w / 2, s.h / 2);
    for (int j = 0; j < d.h; ++j)
        for (int i = 0; i < d.w; ++i) d.at(i, j) = s.at(2 * i, 2 * j);
    return d;
}

struct Cand { double x, y; int z, o; };  // sift.cpp:14

using Octave = std::vector<Plane>;

// sift.cpp:227-256 -- tie-tolerant test over the (2 border + 1)^3 cube (only strict orderings
// disqualify); border = window_size / 2.
bool extremum_cube(const Octave& D, int x, int y, int z, int border) {
    bool mx = true, mn = true;
    double c = D[z].at(x, y);
    for (int dz = -border; dz <= border; ++dz)
        for (int dy = -border; dy <= border; ++dy)
            for (int dx = -border; dx <= border; ++dx) {
                if (!dx && !dy && !dz) continue;
                double n = D[z + dz].at(x + dx, y + dy);
                if (c < n) mx = false;
                if (c > n) mn = false;
            }
    return mx || mn;
}

struct Fit { double off[3]; double g[3]; double H[3][3]; double centre; };

// sift.cpp:32-106 -- cube indexed [z][x][y] of D/255; central differences; closed-form inverse.
Fit fit_cell(const Octave& D, int x, int y, int z) {
    double c[3][3][3];
    for (int dz = 0; dz < 3; ++dz)
        for (int dx = 0; dx < 3; ++dx)
            for (int dy = 0; dy < 3; ++dy)
                c[dz][dx][dy] = D[z + dz - 1].at(x + dx - 1, y + dy - 1) / 255.0;
    Fit f;
    f.centre = c[1][1][1];
    f.g[0] = 0.5 * (c[2][1][1] - c[0][1][1]);
    f.g[1] = 0.5 * (c[1][2][1] - c[1][0][1]);
    f.g[2] = 0.5 * (c[1][1][2] - c[1][1][0]);
    double (*h)[3] = f.H;
    h[0][0] = c[0][1][1] - 2 * c[1][1][1] + c[2][1][1];
    h[1][1] = c[1][0][1] - 2 * c[1][1][1] + c[1][2][1];
    h[2][2] = c[1][1][0] - 2 * c[1][1][1] + c[1][1][2];
    h[0][1] = h[1][0] = 0.25 * (c[2][2][1] - c[2][0][1] - c[0][2][1] + c[0][0][1]);
    h[0][2] = h[2][0] = 0.25 * (c[2][1][2] - c[2][1][0] - c[0][1][2] + c[0][1][0]);
    h[1][2] = h[2][1] = 0.25 * (c[1][0][0] - c[1][2][0] - c[1][0][2] + c[1][2][2]);
    double det = h[0][0] * h[1][1] * h[2][2] + 2 * (h[0][1] * h[1][2] * h[2][0]) -
                 h[0][2] * h[1][1] * h[2][0] - h[0][0] * h[1][2] * h[2][1] -
                 h[0][1] * h[1][0] * h[2][2];
    double i00 = (h[1][1] * h[2][2] - h[1][2] * h[2][1]) / det;
    double i01 = (h[0][2] * h[2][1] - h[0][1] * h[2][2]) / det;
    double i02 = (h[0][1] * h[1][2] - h[0][2] * h[1][1]) / det;
    double i11 = (h[0][0] * h[2][2] - h[0][2] * h[2][0]) / det;
    double i12 = (h[0][2] * h[1][0] - h[0][0] * h[1][2]) / det;
    double i22 = (h[0][0] * h[1][1] - h[0][1] * h[1][0]) / det;
    f.off[0] = -i00 * f.g[0] - i01 * f.g[1] - i02 * f.g[2];
    f.off[1] = -i01 * f.g[0] - i11 * f.g[1] - i12 * f.g[2];
    f.off[2] = -i02 * f.g[0] - i12 * f.g[1] - i22 * f.g[2];
    return f;
}

// sift.hh:31-41 / :25-27
bool kp_less(const OracleKeypoint& a, const OracleKeypoint& b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.size != b.size) return a.size > b.size;
    if (a.pori != b.pori) return a.pori < b.pori;
    return a.octave > b.octave;
}
bool kp_same(const OracleKeypoint& a, const OracleKeypoint& b) {
    return a.x == b.x && a.y == b.y && a.size == b.size && a.pori == b.pori;
}

}  // namespace

struct OracleRun {
    OracleParams p;          // sift.hh:65-71 arguments
    int layers() const { return p.intervals + 3; }  // sift.cpp:144
    int dogs() const { return p.intervals + 2; }    // sift.cpp:212
    std::vector<double> sigmas;
    std::vector<Octave> G, D;
    std::vector<Cand> extrema;
    std::vector<OracleKeypoint> raw, oriented, final_kps;
};

namespace {

// sift.cpp:113-155 + :181-225
void build_pyramid(OracleRun& r, const double* px, int w, int h, int c, bool doubled) {
    const double sigma0 = r.p.init_sigma;
    const int kIntervals = r.p.intervals, kLayers = r.layers(), kDogs = r.dogs();
    Plane base = to_gray(px, w, h, c);
    if (doubled) base = upsample2(base);
    base = blur(base, std::sqrt(sigma0 * sigma0 - 1));  // sift.cpp:124 (always "-1")
    int octaves = (int)std::floor(std::log2(std::min(base.w, base.h) / 3));  // int division
    r.sigmas.assign(kLayers, 0.0);
    r.sigmas[0] = sigma0;
    double k = std::pow(2.0, 1.0 / kIntervals);
    for (int i = 1; i < kLayers; ++i)
        r.sigmas[i] = std::pow(k, i - 1) * sigma0 * std::sqrt(k * k - 1);
    r.G.resize(std::max(octaves, 0));
    r.D.resize(std::max(octaves, 0));
    for (int o = 0; o < octaves; ++o) {
        Octave& g = r.G[o];
        g.resize(kLayers);
        g[0] = base;
        for (int i = 1; i < kLayers; ++i) g[i] = blur(g[i - 1], r.sigmas[i]);  // cascade
        if (g[kLayers - 3].w < 2 || g[kLayers - 3].h < 2) {  // image.cpp:42-44 would throw
            r.G.resize(o + 1);
            r.D.resize(o + 1);
            octaves = o + 1;
        } else {
            base = decimate(g[kLayers - 3]);
        }
    }
    for (int o = 0; o < octaves; ++o) {
        r.D[o].resize(kDogs);
        for (int i = 0; i < kDogs; ++i) {
            const Plane &a = r.G[o][i + 1], &b = r.G[o][i];
            Plane d(a.w, a.h);
            for (size_t t = 0; t < d.v.size(); ++t) d.v[t] = a.v[t] - b.v[t];
            r.D[o][i] = std::move(d);
        }
    }
}

// sift.cpp:264-319 -- x outer, y, z inner; the threshold lands in an int parameter.
void scan_extrema(OracleRun& r) {
    const int kIntervals = r.p.intervals, kDogs = r.dogs();
    const int thr = (int)std::floor(0.5 * r.p.contrast_threshold / (double)kIntervals * 255.0);
    const int border = r.p.window_size / 2;  // sift.cpp:272
    for (int o = 0; o < (int)r.D.size(); ++o) {
        const Octave& D = r.D[o];
        for (int x = border; x < D[0].w - border; ++x)
            for (int y = border; y < D[0].h - border; ++y)
                for (int z = border; z < kDogs - border; ++z) {
                    if (std::abs(D[z].at(x, y)) <= thr) continue;
                    if (extremum_cube(D, x, y, z, border)) r.extrema.push_back({(double)x, (double)y, z, o});
                }
    }
}

// sift.cpp:330-436
void refine(OracleRun& r) {
    const double contrast = r.p.contrast_threshold, ratio = r.p.eigen_ratio, sigma0 = r.p.init_sigma;
    const int kIntervals = r.p.intervals, kDogs = r.dogs();
    const int border = r.p.window_size / 2;  // sift.cpp:336
    for (const Cand& e : r.extrema) {
        const Octave& D = r.D[e.o];
        const int W = D[0].w, H = D[0].h;
        double x = e.x, y = e.y;
        int layer = e.z;
        Fit f;
        bool keep = false;
        for (int step = 0; step < kMaxSteps; ++step) {
            f = fit_cell(D, (int)x, (int)y, layer);
            double m = std::max(std::abs(f.off[0]), std::max(std::abs(f.off[1]), std::abs(f.off[2])));
            if (m < 0.5) {
                double dot = f.g[0] * f.off[0] + f.g[1] * f.off[1] + f.g[2] * f.off[2];
                double val = f.centre + 0.5 * dot;
                if (!((std::abs(val) * kIntervals) >= contrast)) break;
                double tr = f.H[1][1] + f.H[2][2];
                double det = f.H[1][1] * f.H[2][2] - f.H[1][2] * f.H[1][2];
                if (tr <= 0) break;  // Q21: every DoG maximum dies here
                bool edge = (tr * tr * ratio) >= ((ratio + 1) * (ratio + 1) * det);
                keep = !edge;
                break;
            }
            // The reference does `layer += round(off)` on an int with no finiteness guard
            // (sift.cpp:401); a singular Hessian yields inf/NaN whose int conversion on x86-64 is
            // INT_MIN, i.e. out of range -> rejected.  Stated explicitly here.
            if (!std::isfinite(f.off[0]) || !std::isfinite(f.off[1]) || !std::isfinite(f.off[2])) break;
            layer = (int)(layer + std::round(f.off[0]));
            x += std::round(f.off[1]);
            y += std::round(f.off[2]);
            if (x < border || x >= W - border || y < border || y >= H - border || layer < border ||
                layer >= kDogs - border) break;  // sift.cpp:405-410
        }
        if (!keep) continue;
        double s = std::pow(2, e.o);
        OracleKeypoint kp;
        std::memset(&kp, 0, sizeof kp);
        kp.octave = e.o;
        kp.layer = layer;
        kp.x = s * (x + f.off[1]);
        kp.y = s * (y + f.off[2]);
        kp.size = sigma0 * s * std::pow(2, ((double)layer + f.off[0]) / kIntervals);
        kp.pori = 0.0;
        r.raw.push_back(kp);
    }
}

// sift.cpp:447-533
void orient(OracleRun& r, bool doubled) {
    const double peak_ratio = r.p.peak_ratio, factor = r.p.ori_sigma_factor;
    const int kBins = (int)r.p.num_bins;  // declared double, used as int (sift.cpp:450)
    for (const OracleKeypoint& kp : r.raw) {
        double inv = 1.0 / std::pow(2, kp.octave);
        int x = (int)std::round(kp.x * inv), y = (int)std::round(kp.y * inv);
        double scale = factor * (kp.size * inv);
        int radius = (int)std::round(3.0 * scale);
        double denom = 2.0 * scale * scale;
        const Plane& I = r.G[kp.octave][kp.layer];
        std::vector<double> hist_v(kBins, 0.0);
        double* hist = hist_v.data();
        for (int i = -radius; i <= radius; ++i) {
            if (x + i - 1 < 0 || x + i + 1 >= I.w) continue;
            for (int j = -radius; j <= radius; ++j) {
                if (y + j - 1 < 0 || y + j + 1 >= I.h) continue;
                double dx = I.at(x + i + 1, y + j) - I.at(x + i - 1, y + j);
                double dy = I.at(x + i, y + j - 1) - I.at(x + i, y + j + 1);  // up minus down
                double mag = std::sqrt(dx * dx + dy * dy);
                double ang = std::atan2(dy, dx);
                double wgt = std::exp(-(i * i + j * j) / denom);
                int b = (int)std::round(kBins * (ang + kPi) / kTwoPi);  // Q24: bin 0 <-> -pi
                b = (b < kBins) ? b : 0;
                hist[b] += wgt * mag;
            }
        }
        for (int it = 0; it < 2; ++it)  // Q25: in place, sequential
            for (int i = 0; i < kBins; ++i)
                hist[i] = 0.25 * hist[(i - 1 + kBins) % kBins] + 0.5 * hist[i] +
                          0.25 * hist[(i + 1) % kBins];
        double top = *std::max_element(hist, hist + kBins);
        for (int i = 0; i < kBins; ++i) {
            double h0 = hist[(i - 1 + kBins) % kBins], h1 = hist[i], h2 = hist[(i + 1) % kBins];
            if (!(h1 > h0 && h1 > h2 && h1 > (peak_ratio * top))) continue;
            double pos = i + 0.5 * (h0 - h2) / (h0 - 2 * h1 + h2);
            pos = std::fmod(pos + kBins, kBins);
            double ori = kTwoPi * pos / kBins;
            ori = std::fmod(ori + kTwoPi, kTwoPi);
            OracleKeypoint out = kp;
            out.pori = ori;
            if (doubled) { out.x /= 2; out.y /= 2; out.size /= 2; }
            r.oriented.push_back(out);
        }
    }
}

// sift.cpp:541-603 + :610-682
void describe(OracleRun& r, bool doubled) {
    const double scale_factor = r.p.desc_scale_factor;
    for (OracleKeypoint& kp : r.final_kps) {
        const Plane& I = r.G[kp.octave][kp.layer];
        double inv = doubled ? (1.0 / std::pow(2, kp.octave - 1)) : (1.0 / std::pow(2, kp.octave));
        int x = (int)(kp.x * inv), y = (int)(kp.y * inv);  // Q28: truncation
        double size = kp.size * inv;
        double bins_per_rad = 8 / kTwoPi;
        double ca = std::cos(kp.pori), sa = std::sin(kp.pori);
        double hist[128] = {0};
        double hw = scale_factor * size;
        double edenom = 0.5 * 4 * 4;
        double tmp_r = std::round(hw * 0.5 * std::sqrt(2.0) * (4 + 1.0) + 0.5);
        int radius = (int)std::min(tmp_r, std::sqrt(I.w * I.w + I.h * I.h));
        for (int row = -radius; row <= radius; ++row)
            for (int col = -radius; col <= radius; ++col) {
                double rr = (col * sa + row * ca) / hw;
                double cr = (col * ca - row * sa) / hw;
                double rb = rr + 4 / 2 - 0.5, cb = cr + 4 / 2 - 0.5;
                if (!(rb > -1.0 && rb < 4 && cb > -1.0 && cb < 4)) continue;
                int ny = row + y, nx = col + x;
                if (!(nx > 0 && nx < I.w - 1 && ny > 0 && ny < I.h - 1)) continue;
                double dx = I.at(nx + 1, ny) - I.at(nx - 1, ny);
                double dy = I.at(nx, ny - 1) - I.at(nx, ny + 1);
                double mag = std::sqrt(dx * dx + dy * dy);
                double ang = std::atan2(dy, dx);
                ang -= kp.pori;
                ang = std::fmod(std::fmod(ang, kTwoPi) + kTwoPi, kTwoPi);
                double ob = ang * bins_per_rad;
                double wgt = std::exp(-(rr * rr + cr * cr) / edenom);
                double m = mag * wgt;
                // trilinear spread (sift.cpp:541-571)
                int br = (int)std::floor(rb), bc = (int)std::floor(cb), bo = (int)std::floor(ob);
                double fr = rb - br, fc = cb - bc, fo = ob - bo;
                for (int a = 0; a <= 1; ++a) {
                    int ri = br + a;
                    if (ri < 0 || ri >= 4) continue;
                    double vr = m * (a == 0 ? 1.0 - fr : fr);
                    for (int b = 0; b <= 1; ++b) {
                        int ci = bc + b;
                        if (ci < 0 || ci >= 4) continue;
                        double vc = vr * (b == 0 ? 1.0 - fc : fc);
                        for (int d = 0; d <= 1; ++d) {
                            int oi = (bo + d) % 8;
                            hist[(ri * 4 + ci) * 8 + oi] += vc * (d == 0 ? 1.0 - fo : fo);
                        }
                    }
                }
            }
        // sift.cpp:576-603
        double nrm = 0.0;
        for (int i = 0; i < 128; ++i) nrm += hist[i] * hist[i];
        double inv_n = 1.0 / std::sqrt(nrm);
        nrm = 0.0;
        for (int i = 0; i < 128; ++i) {
            hist[i] *= inv_n;
            if (hist[i] > 0.2) hist[i] = 0.2;
            nrm += hist[i] * hist[i];
        }
        inv_n = 1.0 / std::sqrt(nrm);
        for (int i = 0; i < 128; ++i) {
            int q = (int)std::floor(512.0 * hist[i] * inv_n);
            kp.desc[i] = (uint8_t)std::min(q, 255);
        }
    }
}

}  // namespace

extern "C" {

void oracle_default_params(OracleParams* p) {
    p->double_image_size = 1; p->init_sigma = 1.6; p->intervals = 3; p->contrast_threshold = 0.04;
    p->eigen_ratio = 10.0; p->peak_ratio = 0.8; p->ori_sigma_factor = 1.5; p->desc_scale_factor = 3.0;
    p->window_size = 3; p->num_bins = 36;
}

OracleRun* oracle_run_create(const double* pixels, int w, int h, int c, int double_image_size,
                             int keep_pyramid) {
    OracleParams p;
    oracle_default_params(&p);
    p.double_image_size = double_image_size;
    return oracle_run_create_ex(pixels, w, h, c, &p, keep_pyramid);
}

OracleRun* oracle_run_create_ex(const double* pixels, int w, int h, int c, const OracleParams* params,
                                int keep_pyramid) {
    OracleRun* r = new OracleRun();
    r->p = *params;
    bool doubled = params->double_image_size != 0;
    build_pyramid(*r, pixels, w, h, c, doubled);
    scan_extrema(*r);
    refine(*r);
    orient(*r, doubled);
    r->final_kps = r->oriented;
    std::sort(r->final_kps.begin(), r->final_kps.end(), kp_less);  // sift.cpp:20-24
    r->final_kps.erase(std::unique(r->final_kps.begin(), r->final_kps.end(), kp_same),
                       r->final_kps.end());
    describe(*r, doubled);
    if (!keep_pyramid) {
        r->G.clear(); r->G.shrink_to_fit();
        r->D.clear(); r->D.shrink_to_fit();
    }
    return r;
}

void oracle_run_destroy(OracleRun* r) { delete r; }
int oracle_run_octaves(const OracleRun* r) { return (int)r->D.size(); }

int oracle_run_sigmas(const OracleRun* r, double* out, int cap) {
    int n = (int)r->sigmas.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = r->sigmas[i];
    return n;
}

int oracle_run_layer_dims(const OracleRun* r, int octave, int* w, int* h) {
    if (octave < 0 || octave >= (int)r->G.size()) return -1;
    *w = r->G[octave][0].w; *h = r->G[octave][0].h;
    return 0;
}

const double* oracle_run_gaussian(const OracleRun* r, int o, int l) { return r->G[o][l].v.data(); }
const double* oracle_run_dog(const OracleRun* r, int o, int l) { return r->D[o][l].v.data(); }

int oracle_run_extrema(const OracleRun* r, double* out, int cap) {
    int n = (int)r->extrema.size();
    for (int i = 0; i < n && i < cap; ++i) {
        out[4 * i] = r->extrema[i].x; out[4 * i + 1] = r->extrema[i].y;
        out[4 * i + 2] = r->extrema[i].z; out[4 * i + 3] = r->extrema[i].o;
    }
    return n;
}

// Stage functions on CALLER-SUPPLIED keypoints over this run's pyramid (needs keep_pyramid): used by the parity
// tests to attribute a descriptor / orientation difference to the stage that produced it.
// compute_orientations (sift.cpp:447-533) on `n` raw keypoints; returns the number of oriented keypoints.
int oracle_run_orient_given(OracleRun* r, const OracleKeypoint* in, int n, OracleKeypoint* out, int cap) {
    if (r->G.empty()) return -1;
    std::vector<OracleKeypoint> raw(in, in + n), ori;
    raw.swap(r->raw);
    ori.swap(r->oriented);
    orient(*r, r->p.double_image_size != 0);
    const int m = (int)r->oriented.size();
    if (out != nullptr)
        for (int i = 0; i < m && i < cap; ++i) out[i] = r->oriented[i];
    raw.swap(r->raw);
    ori.swap(r->oriented);
    return m;
}

// compute_descriptors (sift.cpp:610-682) on `n` oriented keypoints, in place (fills .desc).
int oracle_run_describe_given(OracleRun* r, OracleKeypoint* inout, int n) {
    if (r->G.empty()) return -1;
    std::vector<OracleKeypoint> kps(inout, inout + n);
    kps.swap(r->final_kps);
    describe(*r, r->p.double_image_size != 0);
    for (int i = 0; i < n; ++i) inout[i] = r->final_kps[i];
    kps.swap(r->final_kps);
    return n;
}

int oracle_run_keypoints(const OracleRun* r, int stage, OracleKeypoint* out, int cap) {
    const std::vector<OracleKeypoint>& v = stage == 0 ? r->raw : stage == 1 ? r->oriented : r->final_kps;
    int n = (int)v.size();
    if (out)
        for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}

// sift.cpp:688-695 + :783-815 -- strict '<' updates, so the lowest j wins ties; emitted in
// ascending i; an empty second set emits nothing, a one-element set always matches.
int oracle_match(const uint8_t* A, int na, const uint8_t* B, int nb, double ratio, int* ia,
                 int* ib, double* dist, int cap) {
    int n = 0;
    for (int i = 0; i < na; ++i) {
        double best = std::numeric_limits<double>::max(), second = best;
        int bj = 0;
        for (int j = 0; j < nb; ++j) {
            double s = 0.0;
            for (int k = 0; k < 128; ++k) {
                int d = (int)A[128 * (size_t)i + k] - (int)B[128 * (size_t)j + k];
                s += d * d;
            }
            double d = std::sqrt(s);
            if (d < best) { second = best; best = d; bj = j; }
            else if (d < second) second = d;
        }
        if (nb > 0 && best < ratio * second) {
            if (n < cap) { ia[n] = i; ib[n] = bj; dist[n] = best; }
            ++n;
        }
    }
    return n;
}

int oracle_gaussian_taps(double sigma, double* taps, int cap) {
    std::vector<double> k = gaussian_taps(sigma);
    for (int i = 0; i < (int)k.size() && i < cap; ++i) taps[i] = k[i];
    return (int)k.size();
}

void oracle_blur(const double* in, int w, int h, double sigma, double* out) {
    Plane p(w, h);
    std::memcpy(p.v.data(), in, sizeof(double) * (size_t)w * h);
    Plane q = blur(p, sigma);
    std::memcpy(out, q.v.data(), sizeof(double) * (size_t)w * h);
}

}  // extern "C"
