// A B200 box as ONE very fast worker of an unmodified MythTracer master.
//
// Speaks the reference's wire protocol (reference VerStarting/network.h:16-26, network.cc:53-133):
//   packet = tag[4] id[8] length_u32 payload;  worker -> master "RDY!" (empty) and "PXLS" (u32 n + RGB24),
//   master -> worker "CAMR" (Camera::Serialize, 56 bytes) and "WORK" (WorkChunk::SerializeInput, 24 bytes),
// and follows the reference worker's loop (main_net_worker.cc:72-168): connect to <master>:12345, introduce
// itself, then render every WORK tile with the last CAMR camera and send the pixels back.  The rendering is
// raytracer::MythTracer from include/mythtracer/ (CUDA; all GPUs of the box can be given with --devices).
//
//   mythtracer_worker_b200 <tag> <master_address> --obj scene.obj [--port 12345] [--depth 5]
//                          [--devices 0,1,2,3] [--lights rig.txt] [--max-chunks N]
//
// The reference worker hard-codes its model path and light rig (main_net_worker.cc:29-65); here the model is
// an argument and the rig defaults to the reference's four lights.
#include <arpa/inet.h>
#include <netdb.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <mythtracer/mythtracer.h>

using raytracer::Camera;
using raytracer::Light;
using raytracer::MythTracer;
using raytracer::WorkChunk;

namespace {

bool ReadAll(int fd, void *buf, size_t n) {
  char *p = static_cast<char *>(buf);
  while (n > 0) {
    const ssize_t r = recv(fd, p, n, 0);
    if (r <= 0) return false;
    p += r;
    n -= (size_t)r;
  }
  return true;
}

bool WriteAll(int fd, const void *buf, size_t n) {
  const char *p = static_cast<const char *>(buf);
  while (n > 0) {
    const ssize_t r = send(fd, p, n, MSG_NOSIGNAL);
    if (r <= 0) return false;
    p += r;
    n -= (size_t)r;
  }
  return true;
}

bool SendPacket(int fd, const char tag[4], const std::string &id, const std::vector<uint8_t> &payload) {
  char id8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  memcpy(id8, id.data(), id.size() < 8 ? id.size() : 8);
  const uint32_t len = (uint32_t)payload.size();
  return WriteAll(fd, tag, 4) && WriteAll(fd, id8, 8) && WriteAll(fd, &len, 4) && (len == 0 || WriteAll(fd, payload.data(), len));
}

bool ReceivePacket(int fd, std::string *tag, std::vector<uint8_t> *payload) {
  uint8_t header[16];
  if (!ReadAll(fd, header, sizeof(header))) return false;
  tag->assign(reinterpret_cast<char *>(header), 4);
  uint32_t len;
  memcpy(&len, header + 12, 4);
  if (len > 1024 * 1024) return false;  // network.cc:101-103
  payload->resize(len);
  return len == 0 || ReadAll(fd, payload->data(), len);
}

int Connect(const char *host, int port) {
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(host, std::to_string(port).c_str(), &hints, &res) != 0) return -1;
  int fd = -1;
  for (addrinfo *a = res; a != nullptr; a = a->ai_next) {
    fd = socket(a->ai_family, a->ai_socktype, a->ai_protocol);
    if (fd < 0) continue;
    if (connect(fd, a->ai_addr, a->ai_addrlen) == 0) break;
    close(fd);
    fd = -1;
  }
  freeaddrinfo(res);
  return fd;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    puts("usage: mythtracer_worker_b200 <tag> <master_address> --obj scene.obj [--port 12345] [--depth 5]\n"
         "       [--devices 0,1,...] [--lights rig.txt] [--max-chunks N]\n"
         "note : tag should have at most 8 characters");
    return 1;
  }
  const std::string id(argv[1]);
  const char *master = argv[2];
  std::string obj, lights_path;
  int port = 12345, depth = raytracer::MAX_RECURSION_LEVEL;
  long max_chunks = -1;
  std::vector<int> devices;
  for (int i = 3; i + 1 < argc; i += 2) {
    const std::string k(argv[i]);
    if (k == "--obj") obj = argv[i + 1];
    else if (k == "--port") port = atoi(argv[i + 1]);
    else if (k == "--depth") depth = atoi(argv[i + 1]);
    else if (k == "--lights") lights_path = argv[i + 1];
    else if (k == "--max-chunks") max_chunks = atol(argv[i + 1]);
    else if (k == "--devices") {
      for (char *tok = strtok(argv[i + 1], ","); tok != nullptr; tok = strtok(nullptr, ",")) devices.push_back(atoi(tok));
    }
  }
  if (obj.empty()) {
    puts("error: --obj is required");
    return 1;
  }

  MythTracer mt;
  if (!devices.empty()) mt.SetDevices(devices);
  mt.SetMaxRecursionLevel(depth);
  if (!mt.LoadObj(obj.c_str())) return 1;
  auto &lights = mt.GetScene()->lights;
  lights.clear();
  if (!lights_path.empty()) {
    FILE *f = fopen(lights_path.c_str(), "r");
    double v[12];
    while (f != nullptr && fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6,
                                  v + 7, v + 8, v + 9, v + 10, v + 11) == 12) {
      lights.push_back(Light{{v[0], v[1], v[2]}, {v[3], v[4], v[5]}, {v[6], v[7], v[8]}, {v[9], v[10], v[11]}});
    }
    if (f != nullptr) fclose(f);
  } else {  // the reference worker's rig (main_net_worker.cc:34-65)
    lights.push_back(Light{{231.82174, 81.69966, -27.78259}, {0.3, 0.3, 0.3}, {1.0, 1.0, 1.0}, {1.0, 1.0, 1.0}});
    for (double z : {0.0, 80.0, 160.0}) lights.push_back(Light{{200, 80.0, z}, {0.0, 0.0, 0.0}, {0.3, 0.3, 0.3}, {0.3, 0.3, 0.3}});
  }
  printf("Name of this worker: %s\n", id.c_str());

  long done = 0;
  for (;;) {
    puts("Connecting...");
    const int fd = Connect(master, port);
    if (fd < 0) {
      printf("error: failed to connect to %s:%d\n", master, port);
      std::this_thread::sleep_for(std::chrono::seconds(1));
      continue;
    }
    puts("Connected!");
    if (!SendPacket(fd, "RDY!", id, {})) {
      puts("error: disconnected when sending RDY!");
      close(fd);
      std::this_thread::sleep_for(std::chrono::seconds(2));
      continue;
    }
    Camera cam{};
    for (;;) {
      std::string tag;
      std::vector<uint8_t> payload;
      if (!ReceivePacket(fd, &tag, &payload)) {
        puts("error: invalid proto or disconnected");
        break;
      }
      if (tag == "CAMR") {
        if (!cam.Deserialize(payload)) {
          puts("error: failed to deserialize camera");
          break;
        }
        continue;
      }
      if (tag == "SCNE") continue;  // an empty stub upstream (network.cc:14-20)
      if (tag != "WORK") {
        puts("error: unexpected packet");
        break;
      }
      WorkChunk work{};
      if (!work.DeserializeInput(payload)) {
        puts("error: failed to deserialize work chunk");
        break;
      }
      work.output_bitmap.resize((size_t)work.chunk_width * work.chunk_height * 3);
      work.camera = cam;
      if (!mt.RayTrace(&work)) {
        puts("error: failed while raytracing; exiting");
        return 1;
      }
      std::vector<uint8_t> out;
      if (!work.SerializeOutput(&out) || !SendPacket(fd, "PXLS", id, out)) {
        puts("error: disconnected when sending PXLS");
        break;
      }
      done++;
      if (max_chunks >= 0 && done >= max_chunks) {
        close(fd);
        printf("Sent %ld chunks; leaving.\n", done);
        return 0;
      }
    }
    close(fd);
    std::this_thread::sleep_for(std::chrono::seconds(2));
  }
}
