"""Cluster types written the way a user of the plugin contract would (README.md:48-88 of the reference):
CUDA source handed to pmdi_register_cluster_type.  USER_GAUSSIAN restates the reference's GaussianCluster
(src/datatypes/gaussian_cluster.jl:12-83) feature by feature, in its operation order; USER_POISSON is a type the
reference does not have (Gamma(1, 1)-Poisson counts), to show that the contract is open."""

USER_GAUSSIAN = r"""
struct MyGaussian {
  static constexpr int WORDS = 4;  // sum, beta, mu, lambda
  __device__ static void init(double* st) { st[0] = 0.0; st[1] = 0.5; st[2] = 0.0; st[3] = 1.0; }
  __device__ static double logprob(const double* st, int n, double x) {
    const double nn = (double)n;
    const double d = x - st[2];
    return log(1.0 / sqrt(3.14159265358979323846)) + lgamma(0.5 * nn + 1.0) - lgamma(0.5 * nn + 0.5)
         + 0.5 * log(st[3] / (nn + 1.0))
         - (0.5 * nn + 1.0) * log(1.0 + (1.0 / (nn + 1.0)) * (d * d) * st[3]);
  }
  __device__ static void add(double* st, int n, double x) {   // n = size after the add
    const double nn = (double)n;
    st[0] = __dadd_rn(st[0], x);
    const double dd = __dadd_rn(x, -st[2]);
    st[1] = __dadd_rn(st[1], __ddiv_rn(__dmul_rn(__dadd_rn(__dadd_rn(nn, -1.0), 0.001), __dmul_rn(dd, dd)),
                                      __dmul_rn(2.0, __dadd_rn(nn, 0.001))));
    st[2] = __ddiv_rn(st[0], __dadd_rn(nn, 0.001));
    st[3] = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(0.5, nn), 0.5), __dadd_rn(nn, 0.001)),
                      __dmul_rn(st[1], __dadd_rn(nn, 1.001)));
  }
  __device__ static double logmarginal(const double* st, int n) {
    const double nn = (double)n, a_n = nn / 2 + 0.5, a_0 = 0.5, b_0 = 0.5, k_0 = 0.001, k_n = nn + k_0;
    return -a_n * log(st[1]) + ((a_0 * log(b_0)) + lgamma(a_n) - lgamma(a_0) + 0.5 * (log(k_0) - log(k_n))
                                - (nn * 0.5) * log(2 * 3.14159265358979323846));
  }
};
"""

USER_POISSON = r"""
struct MyPoisson {   // counts x ~ Poisson(rate), rate ~ Gamma(1, 1): posterior predictive is negative binomial
  static constexpr int WORDS = 1;  // sum of the counts
  __device__ static void init(double* st) { st[0] = 0.0; }
  __device__ static double logprob(const double* st, int n, double x) {
    const double a = 1.0 + st[0], b = 1.0 + (double)n;   // posterior Gamma(a, rate b)
    return lgamma(a + x) - lgamma(a) - lgamma(x + 1.0) + a * log(b / (b + 1.0)) - x * log(b + 1.0);
  }
  __device__ static void add(double* st, int n, double x) { st[0] += x; }
  __device__ static double logmarginal(const double* st, int n) { return lgamma(1.0 + st[0]) - (1.0 + st[0]) * log(1.0 + (double)n); }
};
"""

BROKEN = "struct Broken { static constexpr int WORDS = 2; __device__ static void init(double* st) { st[0] = undefined_symbol; } };"
