// Runs one PPO collect and one evaluate through the C++ host side (include/twisterl_b200.hpp) and dumps the result;
// tests/test_cpp_host.py compares the dump with the Python mirror's collect on the same Philox streams.
//   collect_check <weights.bin> <out.bin> <precision 0..3 = twr_precision> <seed> <collect_id> <episodes> <difficulty>
// weights.bin: float32 [256*512 emb rows][512 emb bias][512*256 W1.T][256 b1][256*4 Wa.T][4 ba][256 Wv.T][1 bv]
#include "twisterl_b200.hpp"

#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    if (argc != 8) { std::fprintf(stderr, "usage\n"); return 2; }
    try {
        std::FILE* f = std::fopen(argv[1], "rb");
        if (!f) return 2;
        auto rd = [&](size_t n) { std::vector<float> v(n); if (std::fread(v.data(), 4, n, f) != n) throw twisterl::Error(-1, "short weights file"); return v; };
        twisterl::PolicyWeights w;
        w.embeddings.vectors.resize(256);
        for (auto& row : w.embeddings.vectors) row = rd(512);
        w.embeddings.bias = rd(512);
        w.embeddings.obs_shape = {256};
        w.common.push_back(twisterl::Linear{rd(512 * 256), rd(256), true});
        w.action_net.push_back(twisterl::Linear{rd(256 * 4), rd(4), false});
        w.value_net.push_back(twisterl::Linear{rd(256), rd(1), false});
        std::fclose(f);

        twisterl::Engine eng(0, static_cast<twr_precision>(std::atoi(argv[3])), std::strtoull(argv[4], nullptr, 0));
        twisterl::Policy policy(eng, w);
        const twr_env_spec env = twisterl::Puzzle(4, 4, std::atoi(argv[7]), 2, 256);
        twisterl::PPOCollector collector((size_t)std::atoll(argv[6]), 0.995f, 0.995f, 32);
        eng.set_collect_id((uint32_t)std::atoi(argv[5]));
        std::unique_ptr<twisterl::Collector> c = std::make_unique<twisterl::PPOCollector>(collector);   // through the trait object
        const twisterl::CollectedData d = c->collect(env, policy);
        eng.set_collect_id(99);
        const auto ev = twisterl::evaluate(env, policy, 64, true, 1);

        // solve from a board two moves away from solved (rl/solve.rs:73-101): the action list must solve it
        twisterl::Envs start(eng, twisterl::Puzzle(4, 4, 1, 2, 256), 1);
        start.set_state({1, 5, 2, 3, 4, 0, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15});
        const auto sol = twisterl::solve(start, policy, false, 64);
        if (sol.first.first == 1.0f) {
            for (size_t act : sol.second) start.step({(int32_t)act});
            if (!start.is_final()[0] || start.reward()[0] != 1.0f) return 5;
        }

        std::FILE* o = std::fopen(argv[2], "wb");
        if (!o) return 2;
        const int64_t R = (int64_t)d.values.size();
        std::fwrite(&R, 8, 1, o);
        for (int64_t r = 0; r < R; ++r) {
            for (size_t i = 0; i < 16; ++i) { const int32_t v = (int32_t)d.obs[r][i]; std::fwrite(&v, 4, 1, o); }
            std::fwrite(d.logits[r].data(), 4, 4, o);
            const int32_t act = (int32_t)d.actions[r], perm = d.perms[r] ? (int32_t)*d.perms[r] : -1;
            std::fwrite(&act, 4, 1, o); std::fwrite(&perm, 4, 1, o);
            std::fwrite(&d.values[r], 4, 1, o); std::fwrite(&d.rewards[r], 4, 1, o);
            std::fwrite(&d.additional_data.at("advs")[r], 4, 1, o); std::fwrite(&d.additional_data.at("rets")[r], 4, 1, o);
        }
        std::fwrite(&ev.first, 4, 1, o); std::fwrite(&ev.second, 4, 1, o);
        std::fclose(o);
        std::printf("records %lld successes %lld eval %.4f %.4f\n", (long long)R, (long long)d.successes, ev.first, ev.second);

        // error behaviour: a value head that does not end in one output is rejected -> twisterl::Error
        twisterl::PolicyWeights bad = w;
        bad.value_net[0] = twisterl::Linear{std::vector<float>(256 * 2, 0.f), std::vector<float>(2, 0.f), false};
        try { twisterl::Policy p2(eng, bad); return 3; } catch (const twisterl::Error& e) { if (e.status != TWR_ERR_UNSUPPORTED) return 4; }
        // a second common layer is fine (general layer stacks)
        twisterl::PolicyWeights deep = w;
        deep.common.push_back(twisterl::Linear{std::vector<float>(256 * 256, 0.f), std::vector<float>(256, 0.f), true});
        twisterl::Policy p3(eng, deep);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
