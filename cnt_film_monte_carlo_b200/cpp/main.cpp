// cpp/main.cpp -- command-line driver, the counterpart of the reference's src/main.cpp (same CLI: argv[1] = input JSON,
// default "input.json"; same "exciton monte carlo" block; same output files), running on libcntmc.so.
//
//   cntmc_main [input.json] [--gpus N] [--steps-per-call N] [--seed S] [--displacements EVERY]
//              [--contacts ITERATIONS [--c1 N] [--c2 N] [--track N]]
//
// --gpus N shards the excitons over GPUs 0..N-1 of the box (one NCCL all-reduce of the ensemble sums per engine call);
// --displacements EVERY also writes particle_dispalcement.{x,y,z}.dat (monte_carlo.cpp:345-380) after every EVERY-th call.
//
// Without --contacts it runs the Green-Kubo loop of src/main.cpp:64-80.  With --contacts it runs ITERATIONS rounds of
// the contact loop of src/main.cpp:98-106 (which the reference never reaches, and which never terminates there).
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>

#include "monte_carlo.hpp"

int main(int argc, char* argv[]) {
  std::time_t start_time = std::time(nullptr);
  std::cout << "\n***\nstart time:\n" << std::asctime(std::localtime(&start_time)) << "***\n\n";

  std::string filename = "input.json";
  long long   steps_per_call = 1024, contact_iterations = -1, c1 = 1100, c2 = 0, n_track = 0, track_max_steps = 1 << 20, n_gpus = 1,
            displacements_every = 0;
  unsigned long long seed = 100;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto value = [&](const char* name) -> long long {
      if (i + 1 >= argc) throw std::invalid_argument(std::string("missing value after ") + name);
      return std::atoll(argv[++i]);
    };
    if (a == "--steps-per-call") steps_per_call = value("--steps-per-call");
    else if (a == "--gpus") n_gpus = value("--gpus");
    else if (a == "--displacements") displacements_every = value("--displacements");
    else if (a == "--seed") seed = (unsigned long long)value("--seed");
    else if (a == "--contacts") contact_iterations = value("--contacts");
    else if (a == "--c1") c1 = value("--c1");
    else if (a == "--c2") c2 = value("--c2");
    else if (a == "--track") n_track = value("--track");
    else if (a == "--track-max-steps") track_max_steps = value("--track-max-steps");
    else filename = a;
  }

  try {
    std::ifstream input_file(filename.c_str());
    if (!input_file) throw std::invalid_argument("cannot open " + filename);
    std::stringstream text;
    text << input_file.rdbuf();
    const cntmc::json::Value j = cntmc::json::parse(text.str());
    if (!j.contains("exciton monte carlo"))
      throw std::invalid_argument("json input file does not contain \"exciton monte carlo\"");
    // if the exciton transfer type is davoody the tubes' description goes with the block (reference main.cpp:51-54)
    cntmc::json::Value json_mc = j.at("exciton monte carlo");
    if (json_mc.at("rate type").as_string() == "davoody") json_mc.obj.push_back({"cnts", j.at("cnts")});
    std::ostringstream block;
    cntmc::json::dump(json_mc, block);

    mc::monte_carlo sim(block.str(), (int)n_gpus);
    sim.set_seed(seed);
    const double time_step = sim.time_step();

    if (contact_iterations < 0) {
      sim.kubo_init();
      sim.save_json_properties();
      sim.kubo_create_particles();
      long long calls = 0;
      while (sim.time() < sim.kubo_max_time()) {
        // how many more steps the reference's `while (time < max)` loop would take, with its own accumulation of _time
        double    t = sim.time();
        long long n = 0;
        while (t < sim.kubo_max_time() && n < steps_per_call) {
          t += time_step;
          ++n;
        }
        sim.kubo_run(time_step, n);
        if (displacements_every > 0 && (++calls % displacements_every) == 0) sim.kubo_save_individual_particle_dispalcements();
        std::cout << "kubo simulation: current time [seconds]: " << std::scientific << sim.time() << " .... "
                  << "max time [seconds]: " << sim.kubo_max_time() << "\r" << std::flush;
      }
      std::cout << std::endl << "Green-Kubo simulation finished!" << std::endl;
      std::cout << "exciton hops: " << sim.hops() << std::endl;
    } else {
      sim.init(c1, c2);
      sim.save_json_properties();
      if (n_track > 0) {  // src/main.cpp:88-92 (100 trajectories there)
        std::cout << "saving particle trajectories ...";
        for (int n = 0; n < (int)n_track; ++n) sim.track_particle(time_step, n, track_max_steps);
        std::cout << "done!" << std::endl;
      }
      std::cout << "\nrunning Monte Carlo:" << std::endl;
      for (long long it = 0; it < contact_iterations; ++it) {
        sim.step(time_step);
        sim.save_metrics(time_step);
        sim.repopulate_contacts();
        std::cout << "simulation time [seconds]: " << std::scientific << sim.time() << " .... "
                  << "number of particles: " << sim.number_of_particles() << "\r" << std::flush;
      }
      std::cout << std::endl;
    }
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }

  std::time_t end_time = std::time(nullptr);
  std::cout << "\nruntime: " << std::difftime(end_time, start_time) << " seconds" << std::endl;
  return 0;
}
