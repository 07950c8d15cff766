// Stand-in for cpp-taskflow 2.4.0 (conanfile.txt:3 of the reference; not
// vendored under /root/reference and not installable offline).
//
// TEST INFRASTRUCTURE.  The reference uses taskflow only as "run these 768
// independent row lambdas on a thread pool" (src/main.cpp:214-236): emplace(),
// Executor::run(), .wait().  No arithmetic lives in the dependency, so this
// shim runs the stored callables under an OpenMP dynamic schedule -- the same
// scheduling the reference's own sandbox uses (sandbox/main.cpp:241).
#ifndef PTB_ORACLE_TASKFLOW_SHIM_HPP
#define PTB_ORACLE_TASKFLOW_SHIM_HPP

#include <cstddef>
#include <functional>
#include <utility>
#include <vector>

namespace tf {

class Taskflow
{
public:
    template<typename F>
    auto emplace(F&& f) -> void
    {
        m_tasks.emplace_back(std::forward<F>(f));
    }

    [[nodiscard]] auto tasks() const noexcept -> std::vector<std::function<void()>> const&
    {
        return m_tasks;
    }

private:
    std::vector<std::function<void()>> m_tasks{};
};

class Executor
{
public:
    struct done
    {
        auto wait() const noexcept -> void
        {
        }
    };

    auto run(Taskflow& flow) -> done
    {
        auto const& tasks = flow.tasks();
        long const n = static_cast<long>(tasks.size());
#pragma omp parallel for schedule(dynamic, 1)
        for(long i = 0; i < n; ++i) {
            tasks[static_cast<std::size_t>(i)]();
        }
        return done{};
    }
};

} // namespace tf

#endif
