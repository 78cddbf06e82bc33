function M = readmatrix(path)
% Minimal readmatrix for Octave versions that lack it: a purely numeric comma-separated file.
  M = dlmread(path, ',');
end
