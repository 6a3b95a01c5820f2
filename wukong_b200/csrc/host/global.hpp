// Global configuration with the reference's key names (core/global.hpp:28-124, core/config.hpp:42-243):
// a file / string of "global_<key> <value>" lines.  Only the keys the hot path reads take effect;
// the others are accepted and stored so an existing Wukong config file loads unchanged.
#pragma once
#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

namespace wukong {

struct Global {
    int num_servers = 1, num_threads = 2, num_proxies = 1, num_engines = 1;
    std::string input_folder;
    int data_port_base = 5500, ctrl_port_base = 9576;
    int rdma_buf_size_mb = 64, rdma_rbf_size_mb = 16;
    bool use_rdma = false;
    int rdma_threshold = 300;
    int mt_threshold = 16;
    bool enable_caching = true, enable_workstealing = false;
    int stealing_pattern = 0;
    bool silent = true;
    bool enable_planner = false;          // only user-defined (.fmt) plans on this path
    bool generate_statistics = false;
    bool enable_vattr = false;
    int memstore_size_gb = 20;
    int est_load_factor = 55;
    int num_gpus = 1;
    int gpu_kvcache_size_gb = 10, gpu_rbuf_size_mb = 32, gpu_rdma_buf_size_mb = 64;
    int gpu_key_blk_size_mb = 16, gpu_value_blk_size_mb = 4;
    bool gpu_enable_pipeline = true;
    std::map<std::string, std::string> unknown;   // keys this build has no use for

    // returns false for a malformed value (reference: ASSERT)
    bool set(const std::string &k, const std::string &v) {
        auto I = [&](int &dst, int lo) { int x = atoi(v.c_str()); if (x < lo) return false; dst = x; return true; };
        auto B = [&](bool &dst) { dst = atoi(v.c_str()) != 0; return true; };
        if (k == "global_num_proxies") return I(num_proxies, 1);
        if (k == "global_num_engines") return I(num_engines, 1);
        if (k == "global_input_folder") {
            if (v.empty()) return false;
            input_folder = v;
            if (input_folder.back() != '/') input_folder += '/';
            return true;
        }
        if (k == "global_data_port_base") return I(data_port_base, 1);
        if (k == "global_ctrl_port_base") return I(ctrl_port_base, 1);
        if (k == "global_memstore_size_gb") return I(memstore_size_gb, 1);
        if (k == "global_est_load_factor") { int x = atoi(v.c_str()); if (x <= 0 || x >= 100) return false; est_load_factor = x; return true; }
        // no RDMA device in this build: the reference then stores 0 whatever the file says (config.hpp:74-85, 93-98) and
        // refuses to switch RDMA on (:120-131); a key the file does not mention keeps its default
        if (k == "global_rdma_buf_size_mb") { rdma_buf_size_mb = 0; return true; }
        if (k == "global_rdma_rbf_size_mb") { rdma_rbf_size_mb = 0; return true; }
        if (k == "global_generate_statistics") return B(generate_statistics);
        if (k == "global_num_gpus") return I(num_gpus, 0);
        if (k == "global_gpu_rdma_buf_size_mb") { gpu_rdma_buf_size_mb = 0; return true; }
        if (k == "global_gpu_rbuf_size_mb") return I(gpu_rbuf_size_mb, 1);
        if (k == "global_gpu_kvcache_size_gb") return I(gpu_kvcache_size_gb, 0);
        if (k == "global_gpu_key_blk_size_mb") return I(gpu_key_blk_size_mb, 1);
        if (k == "global_gpu_value_blk_size_mb") return I(gpu_value_blk_size_mb, 1);
        if (k == "global_use_rdma") { use_rdma = false; return true; }
        if (k == "global_rdma_threshold") return I(rdma_threshold, 0);
        if (k == "global_mt_threshold") return I(mt_threshold, 1);
        if (k == "global_enable_caching") return B(enable_caching);
        if (k == "global_enable_workstealing") return B(enable_workstealing);
        if (k == "global_stealing_pattern") return I(stealing_pattern, 0);
        if (k == "global_silent") return B(silent);
        if (k == "global_enable_planner") return B(enable_planner);
        if (k == "global_enable_vattr") return B(enable_vattr);
        if (k == "global_gpu_enable_pipeline") return B(gpu_enable_pipeline);
        unknown[k] = v;
        return true;
    }
    // "key value" lines; '#' comments and blank lines are skipped (config.hpp:170-201)
    bool load_str(const std::string &text, std::string *bad = nullptr) {
        std::istringstream in(text);
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            std::string k, v;
            if (!(ls >> k) || k[0] == '#') continue;
            ls >> v;
            if (!set(k, v)) { if (bad) *bad = k; return false; }
        }
        num_threads = num_engines + num_proxies;
        return true;
    }
    bool load_file(const std::string &fname, std::string *bad = nullptr) {
        std::ifstream f(fname.c_str());
        if (!f) { if (bad) *bad = fname; return false; }
        std::stringstream ss;
        ss << f.rdbuf();
        return load_str(ss.str(), bad);
    }

    // the reference's own defaults (core/global.hpp:28-124); this mirror deviates in three: enable_planner (no optimiser here),
    // generate_statistics (nothing to collect for it) and num_gpus (the engine IS the GPU path)
    void reference_defaults() {
        use_rdma = true;
        enable_planner = true;
        generate_statistics = true;
        num_gpus = 0;
    }

    // the items reload_config may change while the system runs (set_mutable_config, config.hpp:118-158)
    static bool is_mutable(const std::string &k) {
        static const char *keys[] = {"global_use_rdma", "global_rdma_threshold", "global_mt_threshold", "global_enable_caching",
                                     "global_enable_workstealing", "global_stealing_pattern", "global_silent", "global_enable_planner",
                                     "global_enable_vattr", "global_gpu_enable_pipeline"};
        for (const char *x : keys)
            if (k == x) return true;
        return false;
    }
    void finish(bool gpu_build) {
        num_threads = num_engines + num_proxies + (gpu_build ? num_gpus : 0);      // one agent thread per GPU (config.hpp:213-224)
        mt_threshold = std::max(1, std::min(mt_threshold, num_engines));           // config.hpp:226-227
    }
    // load_config(fname, nsrvs), config.hpp:203-230: a config FILE -- lines starting with '#' and empty lines are skipped, the
    // first two tokens of every other line are key and value, later lines win; unknown keys are reported, not fatal.
    // gpu_build mirrors -DUSE_GPU (num_gpus must then be 1).  Returns false when the file cannot be read or a value is refused.
    bool load_config(const std::string &fname, int nsrvs, bool gpu_build = false, std::string *bad = nullptr) {
        if (nsrvs <= 0) { if (bad) *bad = "nsrvs"; return false; }
        num_servers = nsrvs;
        std::ifstream f(fname.c_str());
        if (!f) { if (bad) *bad = fname; return false; }
        std::map<std::string, std::string> items;   // the reference collects the file into a map first: iteration order = key order
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream ls(line);
            std::string k, v;
            ls >> k >> v;
            items[k] = v;
        }
        for (auto &kv : items)
            if (!set(kv.first, kv.second)) { if (bad) *bad = kv.first; return false; }
        if (gpu_build && num_gpus != 1) { if (bad) *bad = "global_num_gpus"; return false; }
        finish(gpu_build);
        return true;
    }
    // reload_config(str), config.hpp:160-178: whitespace-separated "key value" pairs; only the mutable items are applied
    void reload_config(const std::string &str) {
        std::istringstream iss(str);
        std::string k, v;
        std::map<std::string, std::string> items;
        while (iss >> k >> v) items[k] = v;
        for (auto &kv : items)
            if (is_mutable(kv.first)) set(kv.first, kv.second);
        mt_threshold = std::max(1, std::min(mt_threshold, num_engines));
    }
};

}  // namespace wukong
