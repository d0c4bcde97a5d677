// Microbenchmark of the Go-syntax regexp engine (csrc/host_regex.hpp): g++ -O3 -std=c++17 -o build/rx_bench scripts/regex_bench.cpp && build/rx_bench
#include "../transferia_b200/csrc/host_regex.hpp"
#include <chrono>
#include <cstdio>
#include <random>
int main() {
    std::mt19937 g(1); std::vector<std::string> vals;
    for (int i = 0; i < 20000; i++) { std::string s = "http://"; for (int k = 0; k < 20; k++) s += "abcdefghij."[g() % 11]; s += "/"; s.append(40, 'x'); vals.push_back(s); }
    const char* pats[][2] = {{"^(https?)://([^/]+)", "$2 via $1"}, {"\\d+", "N"}, {"x+", "y"}, {"[_@#&]", "-"}};
    for (auto& pr : pats) {
        tfre::Prog p = tfre::compile(pr[0]); tfre::Template t = tfre::parse_template(pr[1], p); tfre::Machine m(p); std::string out; size_t tot = 0, bytes = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < 5; rep++) for (auto& v : vals) { tfre::replace_all(m, t, (const uint8_t*)v.data(), v.size(), out); tot += out.size(); bytes += v.size(); }
        double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
        printf("%-24s %.0f ns/value  %.1f ns/byte  (insts %zu)\n", pr[0], ns / (5.0 * vals.size()), ns / bytes, p.inst.size());
    }
}
