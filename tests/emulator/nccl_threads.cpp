/**
 *  @file nccl_threads.cpp (tests/emulator)
 *  @brief TEST INFRASTRUCTURE: the nine NCCL entry points the host driver resolves at run time, implemented over host
 *         threads of ONE process, so that the multi-rank exchange of `szs_b200_comm_scores_u64tape` (one broadcast of the
 *         packed inputs, row-sharded scoring, one grouped send/recv gather) runs on the CPU test tier: every "rank" is a
 *         thread with its own engine and scope on the emulated device. Built as `libnccl_emulated.so`; the emulated
 *         library finds it through `SZS_B200_NCCL_LIBRARY`, exactly as the real one finds a real NCCL.
 *
 *  Semantics kept: communicators are created collectively from a shared 128-byte identifier; a broadcast returns on a rank
 *  once its own copy is complete; sends and receives match in order per (source, destination) pair and only make progress
 *  when both sides have been issued — a rank that skips a collective leaves its peers waiting (the tests run with a
 *  deadline for that reason).
 */
#include <nccl.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#define NCCL_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

struct Message {
    void const *data;
    size_t bytes;
    bool taken = false;
};

struct World {
    int size = 0, joined = 0;
    std::mutex mutex;
    std::condition_variable changed;
    // Broadcast rendezvous.
    uint64_t round = 0;
    int arrived = 0;
    void const *root_buffer = nullptr;
    // Point-to-point mailboxes, one queue per (source, destination).
    std::map<std::pair<int, int>, std::deque<std::shared_ptr<Message>>> mailboxes;
};

std::mutex registry_mutex;
std::map<std::string, std::shared_ptr<World>> registry;
unsigned next_identifier = 1;

} // namespace

struct ncclComm {
    std::shared_ptr<World> world;
    int rank = 0;
};

namespace {

struct Pending {
    bool send;
    void *buffer;
    size_t bytes;
    int peer;
    ncclComm *comm;
};
thread_local int group_depth = 0;
thread_local std::vector<Pending> pending;

size_t bytes_of(size_t count, ncclDataType_t type) { return count * (type == ncclUint8 || type == ncclInt8 ? 1 : 4); }

ncclResult_t flush() {
    // Publish every send first, then serve the receives, then wait until the sends were taken: no ordering deadlock
    // between ranks that both send and receive in one group.
    std::vector<std::pair<ncclComm *, std::shared_ptr<Message>>> sent;
    for (Pending const &op : pending)
        if (op.send) {
            World &world = *op.comm->world;
            auto message = std::make_shared<Message>(Message {op.buffer, op.bytes});
            {
                std::lock_guard<std::mutex> lock(world.mutex);
                world.mailboxes[{op.comm->rank, op.peer}].push_back(message);
            }
            world.changed.notify_all();
            sent.emplace_back(op.comm, message);
        }
    for (Pending const &op : pending)
        if (!op.send) {
            World &world = *op.comm->world;
            std::unique_lock<std::mutex> lock(world.mutex);
            auto &queue = world.mailboxes[{op.peer, op.comm->rank}];
            world.changed.wait(lock, [&] { return !queue.empty(); });
            std::shared_ptr<Message> message = queue.front();
            queue.pop_front();
            if (message->bytes != op.bytes) return ncclInternalError; // mismatched send / receive sizes
            std::memcpy(op.buffer, message->data, op.bytes);
            message->taken = true;
            lock.unlock();
            world.changed.notify_all();
        }
    for (auto &entry : sent) {
        World &world = *entry.first->world;
        std::unique_lock<std::mutex> lock(world.mutex);
        world.changed.wait(lock, [&] { return entry.second->taken; });
    }
    pending.clear();
    return ncclSuccess;
}

} // namespace

NCCL_EXPORT ncclResult_t ncclGetUniqueId(ncclUniqueId *identifier) {
    std::lock_guard<std::mutex> lock(registry_mutex);
    std::memset(identifier, 0, sizeof(*identifier));
    std::snprintf(identifier->internal, sizeof(identifier->internal), "emulated-%u", next_identifier++);
    return ncclSuccess;
}

NCCL_EXPORT ncclResult_t ncclCommInitRank(ncclComm_t *comm, int world_size, ncclUniqueId identifier, int rank) {
    std::shared_ptr<World> world;
    {
        std::lock_guard<std::mutex> lock(registry_mutex);
        std::shared_ptr<World> &slot = registry[std::string(identifier.internal, sizeof(identifier.internal))];
        if (!slot) slot = std::make_shared<World>(), slot->size = world_size;
        world = slot;
    }
    if (world->size != world_size || rank < 0 || rank >= world_size) return ncclInternalError;
    {   // Creation is collective: wait for every rank.
        std::unique_lock<std::mutex> lock(world->mutex);
        world->joined++;
        world->changed.notify_all();
        world->changed.wait(lock, [&] { return world->joined >= world->size; });
    }
    *comm = new ncclComm {world, rank};
    return ncclSuccess;
}

NCCL_EXPORT ncclResult_t ncclCommDestroy(ncclComm_t comm) {
    delete comm;
    return ncclSuccess;
}

NCCL_EXPORT ncclResult_t ncclBroadcast(void const *send_buffer, void *receive_buffer, size_t count, ncclDataType_t type,
                                       int root, ncclComm_t comm, cudaStream_t) {
    World &world = *comm->world;
    size_t const bytes = bytes_of(count, type);
    std::unique_lock<std::mutex> lock(world.mutex);
    uint64_t const round = world.round;
    if (comm->rank == root) world.root_buffer = send_buffer;
    // 1. Everybody is here and the root's buffer is known.
    if (++world.arrived == world.size) world.arrived = 0, world.round++, world.changed.notify_all();
    else world.changed.wait(lock, [&] { return world.round != round; });
    void const *const source = world.root_buffer;
    lock.unlock();
    if (comm->rank != root && bytes) std::memcpy(receive_buffer, source, bytes);
    else if (comm->rank == root && receive_buffer != send_buffer && bytes) std::memmove(receive_buffer, send_buffer, bytes);
    // 2. Nobody leaves (and lets the root reuse its buffer) before every copy is done.
    lock.lock();
    uint64_t const second = world.round;
    if (++world.arrived == world.size) world.arrived = 0, world.round++, world.changed.notify_all();
    else world.changed.wait(lock, [&] { return world.round != second; });
    return ncclSuccess;
}

NCCL_EXPORT ncclResult_t ncclGroupStart() {
    group_depth++;
    return ncclSuccess;
}

NCCL_EXPORT ncclResult_t ncclGroupEnd() {
    if (--group_depth > 0) return ncclSuccess;
    group_depth = 0;
    return flush();
}

NCCL_EXPORT ncclResult_t ncclSend(void const *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
    pending.push_back({true, const_cast<void *>(buffer), bytes_of(count, type), peer, comm});
    return group_depth ? ncclSuccess : flush();
}

NCCL_EXPORT ncclResult_t ncclRecv(void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
    pending.push_back({false, buffer, bytes_of(count, type), peer, comm});
    return group_depth ? ncclSuccess : flush();
}

NCCL_EXPORT char const *ncclGetErrorString(ncclResult_t result) { return result == ncclSuccess ? "no error" : "emulated NCCL error"; }
