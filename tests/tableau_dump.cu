// Prints the PRODUCT's coefficient tables (bacon_b200/csrc/tableaux.cuh, adams.cuh: what the kernels are compiled from
// and what the strict kernels' constant memory is filled with) as hexadecimal doubles, one table per line.  Host code
// only; tests/test_abi.py compiles it with nvcc and holds the output against tests/golden/reference_coefficients.json.
#include <cstdio>
#include <cstring>

#include "adams.cuh"
#include "tableaux.cuh"

static void put(const char* name, const double* v, int n) {
    std::printf("%s", name);
    for (int i = 0; i < n; ++i) {
        unsigned long long b;
        std::memcpy(&b, &v[i], 8);
        std::printf(" %016llx", b);
    }
    std::printf("\n");
}

template <class Tab> static void rk(const char* name) {
    constexpr int O = Tab::O;
    char key[64];
    double a[36], c[6], b[6], e[6];
    for (int i = 0; i < O; ++i) {
        c[i] = Tab::c(i);
        b[i] = Tab::b(i);
        e[i] = Tab::e(i);
        for (int j = 0; j < O; ++j) a[i * O + j] = Tab::a(i, j);
    }
    std::snprintf(key, sizeof key, "%s.fast.c", name); put(key, c, O);
    std::snprintf(key, sizeof key, "%s.fast.A", name); put(key, a, O * O);
    std::snprintf(key, sizeof key, "%s.fast.b", name); put(key, b, O);
    std::snprintf(key, sizeof key, "%s.fast.e", name); put(key, e, O);
    const double s = Tab::safety;
    std::snprintf(key, sizeof key, "%s.fast.safety", name); put(key, &s, 1);
    for (int literal = 0; literal < 2; ++literal) {
        bacon::RkTableauRt T;
        bacon::fill_runtime_tableau<Tab>(T, literal != 0);
        for (int i = 0; i < O; ++i)
            for (int j = 0; j < O; ++j) a[i * O + j] = T.a[i][j];
        const char* sem = literal ? "literal" : "corrected";
        std::snprintf(key, sizeof key, "%s.%s.c", name, sem); put(key, T.c, O);
        std::snprintf(key, sizeof key, "%s.%s.A", name, sem); put(key, a, O * O);
        std::snprintf(key, sizeof key, "%s.%s.b", name, sem); put(key, T.b, O);
        std::snprintf(key, sizeof key, "%s.%s.e", name, sem); put(key, T.e, O);
        std::snprintf(key, sizeof key, "%s.%s.safety", name, sem); put(key, &T.safety, 1);
    }
}

template <class C> static void bdf(const char* name) {
    double h[8], l[8];
    for (int i = 0; i < C::O; ++i) {
        h[i] = C::higher(i);
        l[i] = C::lower(i);
    }
    char key[64];
    std::snprintf(key, sizeof key, "%s.higher", name); put(key, h, C::O);
    std::snprintf(key, sizeof key, "%s.lower", name); put(key, l, C::O);
}

template <class C> static void adams(const char* name) {
    double p[8], q[8];
    for (int i = 0; i < C::O; ++i) {
        p[i] = C::predictor(i);
        q[i] = C::corrector(i);
    }
    const double e = C::error;
    char key[64];
    std::snprintf(key, sizeof key, "%s.predictor", name); put(key, p, C::O);
    std::snprintf(key, sizeof key, "%s.corrector", name); put(key, q, C::O);
    std::snprintf(key, sizeof key, "%s.error", name); put(key, &e, 1);
}

int main() {
    rk<bacon::TabRKF45>("RK45");
    rk<bacon::TabBS23>("RK23");
    bdf<bacon::CoefBDF6>("BDF6");
    bdf<bacon::CoefBDF2>("BDF2");
    adams<bacon::CoefAdams5>("Adams5");
    adams<bacon::CoefAdams3>("Adams3");
    return 0;
}
