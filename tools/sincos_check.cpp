// Development aid: accuracy of the table-driven sincos (rc_math.cuh: rc_sincos_tab_core, compiled for the host)
// against long-double libm.   g++ -O2 -I code-robchar_b200/csrc tools/sincos_check.cpp -o /tmp/sincos_check
#include <cstdio>
#include <cmath>
#include <random>
#include "rc_math.cuh"
int main() {
    std::mt19937_64 rng(7);
    double worst = 0, worst_x = 0;
    for (double range : {10.0, 1000.0, 99999.0}) {
        std::uniform_real_distribution<double> u(-range, range);
        double w = 0;
        for (int i = 0; i < 4000000; ++i) {
            const double x = u(rng);
            double s, c;
            rc::rc_sincos_tab_core(x, rc::RC_SC_TAB_HOST, &s, &c);
            const long double sl = sinl((long double)x), cl = cosl((long double)x);
            const double e = fmax(fabs((double)(s - sl)), fabs((double)(c - cl)));
            if (e > w) { w = e; if (e > worst) { worst = e; worst_x = x; } }
        }
        printf("|x| < %-8g max abs error %.3e\n", range, w);
    }
    printf("worst %.3e at x = %.17g\n", worst, worst_x);
    return worst < 2.5e-16 ? 0 : 1;
}
