// GPU power sampling for RunHardware.exe — the counterpart of the reference's PSU power meter
// (powermeter/include/PowerMeter.h:43-90, used at host/RunHardware.cpp:156-185 under MM_POWER_METER):
// a background thread samples the board power every `interval_ms` between Start() and Stop().
// The source is NVML (nvmlDeviceGetPowerUsage), loaded with dlopen so that neither the executables
// nor libmm_b200.so link against the driver library.
#pragma once

#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>

namespace mm {

class PowerMeter {
 public:
  using Sample = std::pair<double, double>;  // (seconds since Start, watts)

  explicit PowerMeter(int interval_ms, unsigned device = 0) : interval_ms_(interval_ms) {
    lib_ = dlopen("libnvidia-ml.so.1", RTLD_NOW);
    if (!lib_) throw std::runtime_error("PowerMeter: libnvidia-ml.so.1 not found");
    auto init = reinterpret_cast<int (*)()>(dlsym(lib_, "nvmlInit_v2"));
    auto handle = reinterpret_cast<int (*)(unsigned, void **)>(dlsym(lib_, "nvmlDeviceGetHandleByIndex_v2"));
    power_ = reinterpret_cast<int (*)(void *, unsigned *)>(dlsym(lib_, "nvmlDeviceGetPowerUsage"));
    if (!init || !handle || !power_ || init() != 0 || handle(device, &device_) != 0) {
      throw std::runtime_error("PowerMeter: NVML initialisation failed");
    }
  }
  ~PowerMeter() {
    Stop();
    if (lib_) {
      auto shutdown = reinterpret_cast<int (*)()>(dlsym(lib_, "nvmlShutdown"));
      if (shutdown) shutdown();
      dlclose(lib_);
    }
  }

  void Start() {
    running_ = true;
    start_ = std::chrono::steady_clock::now();
    thread_ = std::thread([this] {
      while (running_) {
        unsigned milliwatts = 0;
        if (power_(device_, &milliwatts) == 0) {
          const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count();
          samples_.emplace_back(t, 1e-3 * milliwatts);
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(interval_ms_));
      }
    });
  }
  void Stop() {
    running_ = false;
    if (thread_.joinable()) thread_.join();
  }
  std::vector<Sample> const &GetSamples() const { return samples_; }
  double AveragePower() const {
    double sum = 0;
    for (auto const &s : samples_) sum += s.second;
    return samples_.empty() ? 0.0 : sum / static_cast<double>(samples_.size());
  }

 private:
  int interval_ms_;
  void *lib_ = nullptr;
  void *device_ = nullptr;
  int (*power_)(void *, unsigned *) = nullptr;
  std::atomic<bool> running_{false};
  std::chrono::steady_clock::time_point start_;
  std::thread thread_;
  std::vector<Sample> samples_;
};

}  // namespace mm
