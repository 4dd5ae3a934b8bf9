/*
 * b2pt_alias_table.h -- Vose alias-table construction, shared by the product host and the test oracle.
 *
 * Restates `Utils::compute_alias_table` (/root/reference/src/Utils/Utils.cpp:63-131) -- which cannot be compiled
 * from where it lies because Utils.cpp pulls OpenImageDenoise (fetched from the network by the reference's CMake).
 * The tables must be bit-identical to the reference's for the RNG-coupled light selection to match, so the
 * quirks are kept on purpose: double-precision normalisation, FIFO work lists, comparisons against the *float*
 * literal 1.0f, `> 1` to re-queue a large entry, and every left-over entry getting probability 1.
 */
#ifndef B2PT_ALIAS_TABLE_H
#define B2PT_ALIAS_TABLE_H

#include <deque>
#include <vector>

inline void b2pt_build_alias_table(const std::vector<float>& weights, float weights_sum, std::vector<float>& out_probas, std::vector<int>& out_alias)
{
    const size_t n = weights.size();
    if (n == 0)
        return;

    const double total = weights_sum;
    std::vector<double> scaled(n);
    for (size_t i = 0; i < n; i++)
    {
        scaled[i] = static_cast<double>(weights[i]) / total;
        scaled[i] *= n;
    }

    out_probas.resize(n);
    out_alias.resize(n);

    std::deque<int> below, above;
    for (size_t i = 0; i < n; i++)
        (scaled[i] < 1.0f ? below : above).push_back((int)i);

    while (!below.empty() && !above.empty())
    {
        const int s = below.front();
        const int l = above.front();
        below.pop_front();
        above.pop_front();

        out_probas[s] = (float)scaled[s];
        out_alias[s] = l;

        scaled[l] = (scaled[l] + scaled[s]) - 1.0f;
        (scaled[l] > 1.0f ? above : below).push_back(l);
    }

    for (int i : above)
        out_probas[i] = 1.0f;
    for (int i : below)
        out_probas[i] = 1.0f;
}

#endif
