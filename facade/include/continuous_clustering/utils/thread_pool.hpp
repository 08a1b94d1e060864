// The reference's class header pulls in utils/thread_pool.hpp; callers of ContinuousClustering only need it to
// exist. The device pipeline has no stage threads (stream order replaces the job queues), so this is empty.
#ifndef CONTINUOUS_CLUSTERING_THREAD_POOL_HPP
#define CONTINUOUS_CLUSTERING_THREAD_POOL_HPP
#endif
