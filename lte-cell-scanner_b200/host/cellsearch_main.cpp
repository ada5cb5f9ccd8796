// cellsearch_main.cpp - the CellSearch command line (reference src/CellSearch.cpp) on top of the
// B200 drop-in: same flags (-s -e -p -c -r -l -d -i -v -b -h, CellSearch.cpp:117-130), same
// per-centre-frequency loop (:471-569), same result table (:576-614).  Capture buffers come from
// recorded files (-l, capbuf_NNNN.it, src/capbuf.cpp:98-115) or raw rtl_sdr byte dumps
// (capbuf_NNNN.bin, --raw); live rtl-sdr capture is out of scope (no radio, no librtlsdr here).
//
// `CellSearch -l` at the reference's HEAD reads fs_programmed/fc_programmed uninitialised
// (CellSearch.cpp:456-458,480); this program defines them the way LTE-Tracker does:
// fs_programmed = 1.92e6*correction, fc_programmed = fc_requested (LTE-Tracker.cpp:791,609).
#include <getopt.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "it_file_min.hpp"
#include "searcher_dropin.hpp"

using namespace std;
using namespace itpp;

static int verbosity = 1;

static void print_usage() {
  cout << "LTE CellSearch (B200 drop-in) help" << endl << endl;
  cout << "CellSearch -s start_frequency [optional_parameters]" << endl << endl;
  cout << "  -h --help                      print this help screen" << endl;
  cout << "  -v --verbose                   increase status messages from program" << endl;
  cout << "  -b --brief                     reduce status messages from program" << endl;
  cout << "  -s --freq-start fs             frequency where cell search should start" << endl;
  cout << "  -e --freq-end fe               frequency where cell search should end" << endl;
  cout << "  -p --ppm ppm                   crystal remaining PPM error (default 120)" << endl;
  cout << "  -c --correction c              crystal correction factor" << endl;
  cout << "  -l --load                      read captured data from capbuf_XXXX.it files" << endl;
  cout << "  -d --data-dir dir              directory of the capbuf_XXXX.it files" << endl;
  cout << "     --raw                       with -l: read capbuf_XXXX.bin raw rtl_sdr byte dumps instead" << endl;
  cout << "     --sweep                     with -l --raw: all centre frequencies in one batched call (lcs_sweep_search_cu8)" << endl;
  cout << "  -r --record / -i --device-index need a live rtl-sdr dongle: not supported by this build" << endl;
}

static string freq_formatter(const double& freq) {   // CellSearch.cpp:322-341
  stringstream temp;
  if (abs(freq) < 998.0) temp << setw(5) << setprecision(3) << freq << "h";
  else if (abs(freq) < 998000.0) temp << setw(5) << setprecision(3) << freq / 1e3 << "k";
  else if (abs(freq) < 998000000.0) temp << setw(5) << setprecision(3) << freq / 1e6 << "m";
  else if (abs(freq) < 998000000000.0) temp << setw(5) << setprecision(3) << freq / 1e9 << "g";
  else if (abs(freq) < 998000000000000.0) temp << setw(5) << setprecision(3) << freq / 1e12 << "t";
  else temp << freq;
  return temp.str();
}

int main(int argc, char* const argv[]) {
  double freq_start = -1, freq_end = -1, ppm = 120, correction = 1;
  bool save_cap = false, use_recorded_data = false, raw = false, batched = false;
  string data_dir = ".";
  static struct option long_options[] = {
      {"help", no_argument, 0, 'h'},          {"verbose", no_argument, 0, 'v'},       {"brief", no_argument, 0, 'b'},
      {"freq-start", required_argument, 0, 's'}, {"freq-end", required_argument, 0, 'e'}, {"ppm", required_argument, 0, 'p'},
      {"correction", required_argument, 0, 'c'}, {"record", no_argument, 0, 'r'},        {"load", no_argument, 0, 'l'},
      {"data-dir", required_argument, 0, 'd'},   {"device-index", required_argument, 0, 'i'}, {"raw", no_argument, 0, 'R'}, {"sweep", no_argument, 0, 'W'},
      {0, 0, 0, 0}};
  for (;;) {
    int idx = 0;
    int c = getopt_long(argc, argv, "hvbs:e:p:c:rld:i:", long_options, &idx);
    if (c == -1) break;
    char* endp;
    switch (c) {
      case 'h': print_usage(); return -1;
      case 'v': verbosity = 2; break;
      case 'b': verbosity = 0; break;
      case 's': freq_start = strtod(optarg, &endp); if (optarg == endp || *endp) { cerr << "Error: could not parse start frequency" << endl; return -1; } break;
      case 'e': freq_end = strtod(optarg, &endp); if (optarg == endp || *endp) { cerr << "Error: could not parse end frequency" << endl; return -1; } break;
      case 'p': ppm = strtod(optarg, &endp); if (optarg == endp || *endp) { cerr << "Error: could not parse ppm value" << endl; return -1; } break;
      case 'c': correction = strtod(optarg, &endp); if (optarg == endp || *endp) { cerr << "Error: could not parse correction factor" << endl; return -1; } break;
      case 'r': save_cap = true; break;
      case 'l': use_recorded_data = true; break;
      case 'd': data_dir = optarg; break;
      case 'R': raw = true; break;
      case 'W': batched = true; break;
      case 'i': break;
      default: return -1;
    }
  }
  if (optind < argc) { cerr << "Error: unknown/extra arguments specified on command line" << endl; return -1; }
  if (freq_start == -1) { cerr << "Error: must specify a start frequency. (Try --help)" << endl; return -1; }
  if (freq_start < 1e6) { cerr << "Error: start frequency must be greater than 1MHz" << endl; return -1; }
  if (freq_start / 100e3 != std::round(freq_start / 100e3)) {
    freq_start = std::round(freq_start / 100e3) * 100e3;
    cout << "Warning: start frequency has been rounded to the nearest multiple of 100kHz" << endl;
  }
  if (freq_end == -1) freq_end = freq_start;
  if (freq_end < freq_start) { cerr << "Error: end frequency must be >= start frequency" << endl; return -1; }
  if (freq_end / 100e3 != std::round(freq_end / 100e3)) {
    freq_end = std::round(freq_end / 100e3) * 100e3;
    cout << "Warning: end frequency has been rounded to the nearest multiple of 100kHz" << endl;
  }
  if (ppm < 0) { cerr << "Error: ppm value must be positive" << endl; return -1; }
  if (ppm > 200) cout << "Warning: ppm value appears to be set unreasonably high" << endl;
  if (abs(correction - 1) > 1000e-6) cout << "Warning: crystal correction factor appears to be unreasonable" << endl;
  if (save_cap || !use_recorded_data) {
    cerr << "Error: live capture / recording needs an rtl-sdr dongle, which this build does not support; use -l" << endl;
    return -1;
  }
  if (verbosity >= 1) {
    cout << "LTE CellSearch (B200 drop-in, " << lcs_version() << ") beginning" << endl;
    if (freq_start == freq_end) cout << "  Search frequency: " << freq_start / 1e6 << " MHz" << endl;
    else cout << "  Search frequency range: " << freq_start / 1e6 << "-" << freq_end / 1e6 << " MHz" << endl;
    cout << "  PPM: " << ppm << endl;
    stringstream temp;
    temp << setprecision(20) << correction;
    cout << "  correction: " << temp.str() << endl;
    cout << "  Captured data will be read from capbufXXXX." << (raw ? "bin" : "it") << " files" << endl;
  }

  try {
    const double fs_programmed = 1.92e6 * correction;                                   // LTE-Tracker.cpp:791
    const uint16 n_extra = (uint16)floor((freq_start * ppm / 1e6 + 2.5e3) / 5e3);       // CellSearch.cpp:463
    vec f_search_set(2 * n_extra + 1);
    for (int i = 0; i < 2 * n_extra + 1; i++) f_search_set(i) = (i - (int)n_extra) * 5000.0;
    const int n_fc = (int)floor((freq_end - freq_start) / 100e3) + 1;                   // :465
    vector<list<Cell> > detected_cells(n_fc);
    xcorr_pss_skip_debug_outputs(true);
    if (batched) {
      // every centre frequency of the sweep in one call: the raw byte dumps are concatenated and handed to the batched
      // search (same per-channel results as the loop below, CellSearch.cpp:465-558)
      if (!raw) { cerr << "Error: --sweep needs --raw capture files" << endl; return -1; }
      vector<unsigned char> all;
      vector<double> fcs;
      uint32_t n_cap = 0;
      for (int fci = 0; fci < n_fc; fci++) {
        stringstream filename;
        filename << data_dir << "/capbuf_" << setw(4) << setfill('0') << fci << ".bin";
        vector<unsigned char> b;
        if (!lcs_it::read_all(filename.str(), b) || b.size() < 2) { cerr << "Error: cannot read " << filename.str() << endl; return -1; }
        if (fci == 0) n_cap = (uint32_t)(b.size() / 2);
        if (b.size() / 2 != n_cap) { cerr << "Error: capture buffers of a batched sweep must have equal length" << endl; return -1; }
        all.insert(all.end(), b.begin(), b.begin() + (size_t)n_cap * 2);
        fcs.push_back(freq_start + fci * 100e3);
      }
      if (verbosity >= 1) cout << "Examining " << n_fc << " center frequencies in one batched sweep ..." << endl;
      sweep_search_cu8(all, n_cap, fcs, f_search_set, fs_programmed, detected_cells);
      if (verbosity >= 1)
        for (int fci = 0; fci < n_fc; fci++)
          for (list<Cell>::iterator it = detected_cells[fci].begin(); it != detected_cells[fci].end(); ++it) {
            cout << "  Detected a cell!" << endl;
            cout << "    cell ID: " << (*it).n_id_cell() << endl;
            cout << "    RX power level: " << 10 * log10((*it).pss_pow) << " dB" << endl;
            cout << "    residual frequency offset: " << (*it).freq_superfine << " Hz" << endl;
          }
    }
    for (int fci = 0; fci < (batched ? 0 : n_fc); fci++) {
      const double fc_requested = freq_start + fci * 100e3;
      if (verbosity >= 1) cout << "Examining center frequency " << fc_requested / 1e6 << " MHz ..." << endl;
      cvec capbuf;
      const double fc_programmed = fc_requested;                                         // LTE-Tracker.cpp:609
      stringstream filename;
      filename << data_dir << "/capbuf_" << setw(4) << setfill('0') << fci << (raw ? ".bin" : ".it");
      if (verbosity >= 2) cout << "Reading captured data from file: " << filename.str() << endl;
      if (!raw) {
        int fc_file = 0;
        if (!lcs_it::read_capbuf(filename.str(), capbuf, fc_file)) { cerr << "Error: cannot read " << filename.str() << endl; return -1; }
        if (fc_requested != fc_file) {
          cout << "Warning: while reading capture buffer " << fci << ", the read" << endl;
          cout << "center frequency did not match the expected center frequency." << endl;
        }
      } else {
        vector<unsigned char> b;
        if (!lcs_it::read_all(filename.str(), b) || b.size() < 2) { cerr << "Error: cannot read " << filename.str() << endl; return -1; }
        capbuf.set_size((int)(b.size() / 2));
        for (int i = 0; i < capbuf.length(); i++)                                        // capbuf.cpp:172-175
          capbuf(i) = complex<double>((b[2 * i] - 127.0) / 128.0, (b[2 * i + 1] - 127.0) / 128.0);
      }
      const uint8 DS_COMB_ARM = 2;
      mat pow; imat frq; vf3d single, inc; vec sp_incoherent, sp; vcf3d xc; uint16 n_comb_xc, n_comb_sp;
      if (verbosity >= 2) cout << "  Calculating PSS correlations" << endl;
      xcorr_pss(capbuf, f_search_set, DS_COMB_ARM, fc_requested, fc_programmed, fs_programmed, pow, frq, single, inc,
                sp_incoherent, xc, sp, n_comb_xc, n_comb_sp);
      vec Z_th1 = calc_Z_th1(sp_incoherent, n_comb_xc, DS_COMB_ARM);                     // :500-503
      if (verbosity >= 2) cout << "  Searching for and examining correlation peaks..." << endl;
      list<Cell> peaks;
      peak_search(pow, frq, Z_th1, f_search_set, fc_requested, fc_programmed, single, DS_COMB_ARM, peaks);
      detected_cells[fci] = peaks;
      list<Cell>::iterator it = detected_cells[fci].begin();
      while (it != detected_cells[fci].end()) {
        vec a, b2; cvec c1, c2, c3, c4; mat l1, l2;
        (*it) = sss_detect((*it), capbuf, 3, fc_requested, fc_programmed, fs_programmed, a, b2, c1, c2, c3, c4, l1, l2);
        if ((*it).n_id_1 == -1) { it = detected_cells[fci].erase(it); continue; }
        (*it) = pss_sss_foe((*it), capbuf, fc_requested, fc_programmed, fs_programmed);
        cmat tfg, tfg_comp; vec ts, ts_comp;
        extract_tfg((*it), capbuf, fc_requested, fc_programmed, fs_programmed, tfg, ts);
        RS_DL rs_dl((*it).n_id_cell(), 6, (*it).cp_type);
        (*it) = tfoec((*it), tfg, ts, fc_requested, fc_programmed, rs_dl, tfg_comp, ts_comp);
        (*it) = decode_mib((*it), tfg_comp, rs_dl);
        if ((*it).n_rb_dl == -1) { it = detected_cells[fci].erase(it); continue; }
        if (verbosity >= 1) {
          cout << "  Detected a cell!" << endl;
          cout << "    cell ID: " << (*it).n_id_cell() << endl;
          cout << "    RX power level: " << 10 * log10((*it).pss_pow) << " dB" << endl;
          cout << "    residual frequency offset: " << (*it).freq_superfine << " Hz" << endl;
        }
        ++it;
      }
    }
    list<Cell> cells_final;
    dedup(detected_cells, cells_final);
    if (cells_final.size() == 0) {
      cout << "No LTE cells were found..." << endl;
    } else {   // CellSearch.cpp:579-613
      cout << "Detected the following cells:" << endl;
      cout << "A: #antenna ports C: CP type ; P: PHICH duration ; PR: PHICH resource type" << endl;
      cout << "CID A      fc   foff RXPWR C nRB P  PR CrystalCorrectionFactor" << endl;
      for (list<Cell>::iterator it = cells_final.begin(); it != cells_final.end(); ++it) {
        stringstream ss;
        ss << setw(3) << (*it).n_id_cell();
        ss << setw(2) << (int)(*it).n_ports;
        ss << " " << setw(6) << setprecision(5) << (*it).fc_requested / 1e6 << "M";
        ss << " " << freq_formatter((*it).freq_superfine);
        ss << " " << setw(5) << setprecision(3) << 10 * log10((*it).pss_pow);
        ss << " " << (((*it).cp_type == cp_type_t::NORMAL) ? "N" : (((*it).cp_type == cp_type_t::UNKNOWN) ? "U" : "E"));
        ss << " " << setw(3) << (int)(*it).n_rb_dl;
        ss << " " << (((*it).phich_duration == phich_duration_t::NORMAL) ? "N" : (((*it).phich_duration == phich_duration_t::UNKNOWN) ? "U" : "E"));
        switch ((*it).phich_resource) {
          case phich_resource_t::UNKNOWN: ss << " UNK"; break;
          case phich_resource_t::oneSixth: ss << " 1/6"; break;
          case phich_resource_t::half: ss << " 1/2"; break;
          case phich_resource_t::one: ss << " one"; break;
          case phich_resource_t::two: ss << " two"; break;
        }
        const double crystal_freq_actual = (*it).fc_requested - (*it).freq_superfine;
        const double correction_new = correction * ((*it).fc_requested / crystal_freq_actual);
        ss << " " << setprecision(20) << correction_new;
        cout << ss.str() << endl;
      }
    }
  } catch (const char* msg) {
    cerr << "Error: " << msg << endl;
    return -1;
  }
  return 0;
}
