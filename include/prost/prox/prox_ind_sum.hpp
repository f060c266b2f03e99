// prost/prox/prox_ind_sum.hpp -- ProxIndSum<T>: index-list groups projected onto a prescribed sum in the metric
// of the step sizes (reference: include/prost/prox/prox_ind_sum.hpp:37-62, src/prox/prox_ind_sum.cu:33-145).
#ifndef PROST_PROX_IND_SUM_HPP_
#define PROST_PROX_IND_SUM_HPP_

#include "prost/prox/prox.hpp"

namespace prost {

template <typename T>
class ProxIndSum : public Prox<T> {
 public:
  ProxIndSum(size_t index, size_t size, size_t count, size_t dim, std::vector<size_t>& inds, T sum)
      : Prox<T>(index, size, true), count_(count), dim_(dim), count_2_(0), dim_2_(0), inds_(inds.begin(), inds.end()),
        sum_(sum), sum_2_(0), two_(false) {}
  ProxIndSum(size_t index, size_t size, size_t count, size_t dim, std::vector<size_t>& inds, T sum, size_t count2,
             size_t dim2, std::vector<size_t>& inds2, T sum2)
      : Prox<T>(index, size, true), count_(count), dim_(dim), count_2_(count2), dim_2_(dim2),
        inds_(inds.begin(), inds.end()), inds_2_(inds2.begin(), inds2.end()), sum_(sum), sum_2_(sum2), two_(true) {}

 protected:
  virtual pb_prox* create() {
    // ProxIndSum::Initialize (prox_ind_sum.cu:72-86)
    if (count_ * dim_ != inds_.size() || (two_ && count_2_ * dim_2_ != inds_2_.size()))
      throw Exception("ProxIndSum: dimensions dont fit");
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_ind_sum_indexed(detail::context(), this->index_, this->size_, count_, dim_,
                                                 inds_.data(), sum_, count_2_, dim_2_,
                                                 two_ ? inds_2_.data() : nullptr, sum_2_, &h));
    return h;
  }
  size_t count_, dim_, count_2_, dim_2_;
  std::vector<unsigned long long> inds_, inds_2_;
  float sum_, sum_2_;
  bool two_;
};

}  // namespace prost

#endif
