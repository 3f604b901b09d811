// Tool (not product): fp64 [F][A][3] binary (Angstrom) -> .xyz text, multi-threaded.
//   xyz_writer in.bin nat nframes symbols.txt out.xyz
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
    if (argc != 6) return 2;
    const long nat = atol(argv[2]), F = atol(argv[3]);
    std::vector<std::string> sym(nat);
    {
        FILE* f = fopen(argv[4], "r");
        char buf[16];
        for (long a = 0; a < nat; ++a) { if (fscanf(f, "%15s", buf) != 1) return 3; sym[a] = buf; }
        fclose(f);
    }
    std::vector<double> x((size_t)F * nat * 3);
    {
        FILE* f = fopen(argv[1], "rb");
        if (!f || fread(x.data(), 8, x.size(), f) != x.size()) return 4;
        fclose(f);
    }
    const int nt = (int)std::thread::hardware_concurrency() > 0 ? (int)std::thread::hardware_concurrency() : 4;
    const long chunk = 256;
    FILE* out = fopen(argv[5], "w");
    for (long f0 = 0; f0 < F; f0 += chunk * nt) {
        std::vector<std::string> parts(nt);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t)
            th.emplace_back([&, t]() {
                const long a0 = f0 + t * chunk, a1 = a0 + chunk < F ? a0 + chunk : F;
                std::string& s = parts[t];
                char line[128];
                for (long f = a0; f < a1; ++f) {
                    s += std::to_string(nat) + "\nframe " + std::to_string(f + 1) + "\n";
                    for (long a = 0; a < nat; ++a) {
                        const double* p = &x[((size_t)f * nat + a) * 3];
                        int n = snprintf(line, sizeof(line), "%s %.10f %.10f %.10f\n", sym[a].c_str(), p[0], p[1], p[2]);
                        s.append(line, n);
                    }
                }
            });
        for (auto& t : th) t.join();
        for (auto& s : parts) fwrite(s.data(), 1, s.size(), out);
    }
    fclose(out);
    return 0;
}
