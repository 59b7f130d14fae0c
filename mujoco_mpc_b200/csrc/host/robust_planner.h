// robust_planner.h - C++ host side of the Robust planner (mjpc/planners/robust/robust_planner.h:30-76,
// robust_planner.cc:40-160) over the SamplingPlanner delegate.  The ncandidates x nrepetitions NoisyRollouts the
// reference schedules on its ThreadPool (robust_planner.cc:110-129) are ONE mjpc_b200_rollout_spline launch with the
// handle's force noise switched on (mjpc_b200_set_xfrc_noise; injected Philox stream, seed + iteration).
#pragma once
#include <memory>

#include "sampling_planner.h"

namespace mjpc_b200_host {

class RobustPlanner {
 public:
  // `noisy` is a second engine handle for the perturbed rollouts, so the delegate's clean trajectories (what
  // BestTrajectory returns, robust_planner.cc:163-165) survive the second launch; owned by this object
  RobustPlanner(std::unique_ptr<SamplingPlanner> delegate, mjpc_b200_t* noisy) : delegate_(std::move(delegate)), noisy_(noisy) {}
  ~RobustPlanner() { if (noisy_) mjpc_b200_destroy(noisy_); }
  // robust_repetitions (5), robust_candidates (-1 -> sampling_trajectories / repetitions), robust_xfrc (0.1),
  // robust_xfrc_rate (0.1): robust_planner.cc:44-57
  void Configure(int sampling_trajectories, int ncandidates, int nrepetitions, double xfrc_std, double xfrc_rate, uint32_t seed);
  void Reset(int horizon, const double* initial_repeated_action) { delegate_->Reset(horizon, initial_repeated_action); }
  void SetState(const double* state, double time, const double* mocap);
  int OptimizePolicy(int horizon);                                     // :91-157
  void ActionFromPolicy(double* action, double time, bool use_previous = false) {
    delegate_->ActionFromPolicy(action, time, use_previous);
  }
  const Trajectory* BestTrajectory() { return delegate_->BestTrajectory(); }
  SamplingPlanner* delegate() { return delegate_.get(); }
  mjpc_b200_t* noisy() { return noisy_; }
  const std::vector<double>& scores() const { return scores_; }

 private:
  std::unique_ptr<SamplingPlanner> delegate_;
  mjpc_b200_t* noisy_ = nullptr;
  int ncandidates_ = 12, nrepetitions_ = 5;
  double xfrc_std_ = 0.1, xfrc_rate_ = 0.1;
  uint32_t seed_ = 0x5EED;
  std::vector<double> state_, mocap_, scores_;
  double time_ = 0;
};

}  // namespace mjpc_b200_host
